"""The block schedule behind mb_iterate's hidden copies (mb_api.cu: iterate_wavefront), checked on the CPU through the
library's own arithmetic (mb_wavefront_*): simulate what every launch reads and writes and make sure that
  * a pass only ever reads planes that hold exactly the previous pass's result (or arrived input for pass 0),
  * overwriting a plane never destroys a value some later launch still needs,
  * no plane is uploaded into after a launch has written there, and planes sent home early are final,
  * every pass ends up covering every plane exactly once."""
import ctypes as C

import pytest

from massiv_b200 import _abi

HALO = 2  # planes a pass reads either side of what it writes (two fused steps of a 3-wide window)


def schedule(L, D, nb, npass):
    """yield ('upload', z0, z1) / ('pass', q, o0, o1) / ('send', z0, z1) in issue order"""
    sent = 0
    for k in range(nb):
        z0 = L.mb_wavefront_arrived(D, nb, k - 1)
        yield ("upload", z0, L.mb_wavefront_arrived(D, nb, k))
        for q in range(npass):
            o0, o1 = C.c_int64(), C.c_int64()
            L.mb_wavefront_region(D, nb, k, q, C.byref(o0), C.byref(o1))
            if o1.value > o0.value:
                yield ("pass", q, o0.value, o1.value)
        done = L.mb_wavefront_final(D, nb, k, npass)
        if done > sent:
            yield ("send", sent, done)
            sent = done


@pytest.mark.parametrize("D,nb,npass", [(1536, 8, 250), (64, 8, 3), (16, 8, 2), (67, 8, 9), (1535, 8, 1000), (40, 4, 7), (129, 16, 30), (9, 3, 1)])
def test_schedule_respects_every_dependency(D, nb, npass):
    L = _abi.lib()
    # level[b][z]: which pass's result plane z of buffer b holds (0 = input, -1 = nothing yet); pass q reads buffer q & 1
    level = [[-1] * D, [-1] * D]
    covered = [[0] * D for _ in range(npass)]
    uploads_done = 0
    for ev in schedule(L, D, nb, npass):
        if ev[0] == "upload":
            _, z0, z1 = ev
            assert z0 == uploads_done and z1 >= z0
            for z in range(z0, z1):
                assert level[0][z] == -1, "an upload must not land on a plane a pass has already written"
                level[0][z] = 0
            uploads_done = z1
        elif ev[0] == "pass":
            _, q, o0, o1 = ev
            src, dst = level[q & 1], level[1 - (q & 1)]
            for z in range(max(0, o0 - HALO), min(D, o1 + HALO)):
                assert src[z] == q, (D, nb, npass, q, o0, o1, z, src[z])   # exactly the previous pass's plane (planes outside the array are the border)
            for z in range(o0, o1):
                # what is overwritten (pass q - 1's INPUT, level q - 1) must have no reader left: every pass q - 1 output
                # that reads it lies within HALO planes and must already exist — they are this launch's own inputs
                assert dst[z] in (-1, q - 1), (q, z, dst[z])
                dst[z] = q + 1
                covered[q][z] += 1
        else:
            _, z0, z1 = ev
            fin = level[npass & 1]
            assert all(fin[z] == npass for z in range(z0, z1)), "a plane was sent home before its last pass"
    assert uploads_done == D
    assert all(c == 1 for row in covered for c in row), "every pass must produce every plane exactly once"
    assert all(v == npass for v in level[npass & 1])


def test_later_launches_never_touch_planes_already_sent():
    L = _abi.lib()
    D, nb, npass = 1536, 8, 250
    sent_upto = 0
    for ev in schedule(L, D, nb, npass):
        if ev[0] == "send":
            sent_upto = ev[2]
        elif ev[0] == "pass" and (1 - (ev[1] & 1)) == (npass & 1):
            assert ev[2] >= sent_upto, "a launch writes the result array below the planes already on their way home"
