"""Device-side helpers the bench's verify block relies on, and iteration with a shifted (non-mapStencil)
size-preserving padding (the single-GPU side of the sharded "sum shifted" case)."""
import ctypes as C

import numpy as np
import pytest

import massiv_b200 as M
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_compare_dev_counts_and_locates_differences():
    dev = M.Device.default()
    rng = np.random.default_rng(1)
    for nbytes in (0, 1, 7, 8, 4097, 1 << 20):
        a = rng.integers(0, 256, nbytes, dtype=np.uint8)
        b = a.copy()
        da, db = dev.fromHost(a), dev.fromHost(b)
        n, first = C.c_uint64(), C.c_uint64()
        dev.check(dev._L.mb_compare_dev(dev.handle, C.c_void_p(da.ptr), C.c_void_p(db.ptr), nbytes, C.byref(n), C.byref(first)))
        assert n.value == 0 and first.value == 2 ** 64 - 1
        if nbytes < 8:
            continue
        where = int(rng.integers(0, nbytes))
        b[where] ^= 0x40
        db = dev.fromHost(b)
        dev.check(dev._L.mb_compare_dev(dev.handle, C.c_void_p(da.ptr), C.c_void_p(db.ptr), nbytes, C.byref(n), C.byref(first)))
        assert n.value == 1 and first.value in (where, where // 8 * 8)


@pytest.mark.parametrize("border", ["fill", "wrap", "reflect"])
def test_iterate_with_shifted_padding(border):
    rng = np.random.default_rng(3)
    arr = rng.integers(-99, 99, (37, 200)).astype(np.int64)
    b = {"fill": M.Fill(0), "wrap": M.Wrap, "reflect": M.Reflect}[border]
    pad = M.Padding((2, 0), (0, 2), b)
    got = M.iterateStencil(4, pad, M.sumStencil(M.Sz(3, 3)), arr)
    want = O.iterate("sum", arr, 4, size=(3, 3), pad_lo=(2, 0), pad_hi=(0, 2), border=border, fill=0)
    assert got.tobytes() == want.tobytes()
