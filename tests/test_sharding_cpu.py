"""Multi-rank host logic on CPU: the slab plan of the C library (`mb_shard_plan_make`, no GPU needed)
and the halo-exchange choreography, run with world_size 2 and 3 over the gloo backend with the CPU
oracle standing in for the per-slab kernel.  The result must equal the oracle's single-process
iteration of the global array, for every border."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import massiv_b200 as M
from massiv_b200 import _abi, sharding
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def oracle_compute(ld, slab):
    """one step of the slab problem through the oracle (same descriptor layout on both sides)"""
    od = O._Desc.from_buffer_copy(bytes(ld))
    dims = O._dims(slab.shape)
    out_dims = (C.c_int64 * O.MAX_RANK)()
    assert O.lib().mo_out_dims(C.byref(od), dims, out_dims) == 0
    out = np.empty(tuple(out_dims[i] for i in range(slab.ndim)), slab.dtype)
    slab = np.ascontiguousarray(slab)
    assert O.lib().mo_apply(C.byref(od), dims, slab.ctypes.data, out.ctypes.data, 1) == 0
    return out


CASES = {
    "heat3d-fill0": dict(shape=(13, 6, 7), dtype=np.float64, border=("fill", 0.0), kind="conv", size=(3, 3, 3)),
    "heat3d-fill2": dict(shape=(12, 5, 6), dtype=np.float64, border=("fill", 2.5), kind="conv", size=(3, 3, 3)),
    "heat3d-wrap": dict(shape=(12, 5, 6), dtype=np.float64, border=("wrap",), kind="conv", size=(3, 3, 3)),
    "heat3d-edge": dict(shape=(11, 5, 6), dtype=np.float64, border=("edge",), kind="conv", size=(3, 3, 3)),
    "heat3d-reflect": dict(shape=(11, 5, 6), dtype=np.float64, border=("reflect",), kind="conv", size=(3, 3, 3)),
    "heat3d-continue": dict(shape=(11, 5, 6), dtype=np.float64, border=("continue",), kind="conv", size=(3, 3, 3)),
    # star-shaped + finite promise + Fill 0: eligible for two steps per pass, so slabs carry TWO ghost planes
    "heat3d-star-tb2": dict(shape=(14, 6, 8), dtype=np.float64, border=("fill", 0.0), kind="conv", size=(3, 3, 3), star=True),
    # the same without the promise (two ghost planes as well: the device runs checked two-step rounds), with Inf / NaN
    # on and next to slab boundaries: the strict 27-tap semantics must survive the sharding
    "heat3d-star-strict-nonfinite": dict(shape=(14, 6, 8), dtype=np.float64, border=("fill", 0.0), kind="conv", size=(3, 3, 3),
                                         star=True, promise=False, nonfinite=True),
    "life-wrap": dict(shape=(16, 12), dtype=np.uint8, border=("wrap",), kind="life", size=(3, 3)),
    "sum5-reflect": dict(shape=(17, 9), dtype=np.int64, border=("reflect",), kind="sum", size=(5, 2)),
    "corr4-edge": dict(shape=(19, 8), dtype=np.float64, border=("edge",), kind="corr", size=(4, 3)),
    # size-preserving descriptors that are NOT mapStencil-shaped along dim 0 (pad_lo != centre): the halo is the
    # padding, not centre / size-1-centre.  sumStencil (3,1) with Padding (2,0) (0,0): window = planes o-2 .. o
    "sum3-shifted-fill": dict(shape=(13, 5), dtype=np.int64, border=("fill", 0), kind="sum", size=(3, 1), pad=((2, 0), (0, 0))),
    "sum3-shifted-wrap": dict(shape=(12, 5), dtype=np.int64, border=("wrap",), kind="sum", size=(3, 1), pad=((1, 0), (1, 0))),
    "corr3-shifted-reflect": dict(shape=(14, 6), dtype=np.float64, border=("reflect",), kind="corr", size=(3, 3), pad=((0, 1), (2, 1))),
}


def make_case(name):
    c = CASES[name]
    import zlib
    rng = np.random.default_rng(zlib.crc32(name.encode()))  # identical in every spawned rank
    if c["dtype"] == np.uint8:
        arr = (rng.random(c["shape"]) < 0.4).astype(np.uint8)
    elif c["dtype"] == np.int64:
        arr = rng.integers(-9, 9, c["shape"]).astype(np.int64)
    else:
        arr = rng.random(c["shape"])
    n = int(np.prod(c["size"]))
    coeffs = rng.standard_normal(n) if c["kind"] in ("conv", "corr") else None
    flags = 0
    if c.get("star"):
        coeffs = np.where(np.isin(np.arange(27), (4, 10, 12, 13, 14, 16, 22)), coeffs, 0.0)
        flags = _abi.FLAG_ASSUME_FINITE if c.get("promise", True) else 0
    if c.get("nonfinite"):
        arr[4, 2, 3], arr[5, 0, 0], arr[6, 5, 7], arr[9, 3, 3] = np.inf, np.nan, -np.inf, np.inf
    b = c["border"]
    border = M.Fill(b[1]) if b[0] == "fill" else {"wrap": M.Wrap, "edge": M.Edge, "reflect": M.Reflect, "continue": M.Continue}[b[0]]
    st = M.Stencil(c["kind"], tuple(c["size"]), None, None if coeffs is None else coeffs.reshape(c["size"]), flags=flags)
    okw = dict(size=c["size"], border=b[0], fill=b[1] if b[0] == "fill" else 0,
               coeffs=None if coeffs is None else coeffs.tolist())
    if "pad" in c:
        okw["pad_lo"], okw["pad_hi"] = c["pad"]
    return arr, border, st, c["kind"], okw


def make_dw(border, st, arr, okw):
    if "pad_lo" in okw:
        return M.applyStencil(M.Padding(tuple(okw["pad_lo"]), tuple(okw["pad_hi"]), border), st, arr)
    return M.mapStencil(border, st, arr)


def _worker(rank, world, port, names, steps, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for name in names:
            arr, border, st, kind, okw = make_case(name)
            desc = make_dw(border, st, arr, okw).desc()
            sp = sharding.shard_plan(desc, arr.shape, world, rank)
            local = arr[sp.first_plane:sp.first_plane + sp.local_planes]
            out = sharding.iterate_sharded_hostcheck(steps, desc, local, arr.shape, dist, oracle_compute,
                                                     fill_value=okw["fill"])
            q.put((name, rank, sp.first_plane, out))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_choreography_matches_global_oracle(world):
    """one gloo job per world size runs every case (all five borders, ranks 2 and 3, asymmetric halos)"""
    steps = 3
    names = sorted(CASES)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, names, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world * len(names))]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for name in names:
        parts = sorted([t for t in got if t[0] == name], key=lambda t: t[2])
        assert len(parts) == world
        res = np.concatenate([t[3] for t in parts], axis=0)
        arr, border, st, kind, okw = make_case(name)
        want = O.iterate(kind, arr, steps, **okw)
        assert res.shape == want.shape, name
        nan = np.isnan(want) if want.dtype.kind == "f" else np.zeros(want.shape, bool)
        assert np.array_equal(np.isnan(res) if res.dtype.kind == "f" else nan, nan), name
        assert res[~nan].tobytes() == want[~nan].tobytes(), name


def test_single_rank_hostcheck_covers_every_border():
    for name in sorted(CASES):
        arr, border, st, kind, okw = make_case(name)
        desc = make_dw(border, st, arr, okw).desc()
        got = sharding.iterate_sharded_hostcheck(2, desc, arr, arr.shape, None, oracle_compute, fill_value=okw["fill"])
        assert got.tobytes() == O.iterate(kind, arr, 2, **okw).tobytes(), name


def test_shard_plan_partition_and_ghost_sources():
    k = np.zeros((3, 3, 3))
    k[0, 0, 0] = 1.0   # a corner tap: not star-shaped, so one step per round
    arr = np.zeros((10, 4, 4))
    st = M.makeConvolutionStencilFromKernel(k)
    # Fill: physical borders are constant planes, inner boundaries come from the neighbours
    d = M.mapStencil(M.Fill(0.0), st, arr).desc()
    plans = [sharding.shard_plan(d, arr.shape, 4, r) for r in range(4)]
    assert [p.first_plane for p in plans] == [0, 2, 5, 7] and [p.local_planes for p in plans] == [2, 3, 2, 3]
    assert all(p.halo_lo == 1 and p.halo_hi == 1 for p in plans)
    assert plans[0].recv_lower_from == -1 and plans[0].ghost_lo_src[0] == _abi.GHOST_FILL and plans[0].recv_upper_from == 1
    assert plans[0].send_first_to == -1 and plans[0].send_last_to == 1
    assert plans[3].recv_upper_from == -1 and plans[3].ghost_hi_src[0] == _abi.GHOST_FILL and plans[3].recv_lower_from == 2
    assert plans[1].recv_lower_from == 0 and plans[1].recv_upper_from == 2 and plans[1].ghost_lo_src[0] == _abi.GHOST_REMOTE
    assert plans[1].send_first_to == 0 and plans[1].send_last_to == 2
    # Wrap: a ring
    d = M.mapStencil(M.Wrap, st, arr).desc()
    p0, p3 = sharding.shard_plan(d, arr.shape, 4, 0), sharding.shard_plan(d, arr.shape, 4, 3)
    assert p0.recv_lower_from == 3 and p3.recv_upper_from == 0 and p0.send_first_to == 3 and p3.send_last_to == 0
    one = sharding.shard_plan(d, arr.shape, 1, 0)
    assert one.recv_lower_from == -1 and one.recv_upper_from == -1 and one.ghost_lo_src[0] == 9 and one.ghost_hi_src[0] == 0
    # Edge / Reflect / Continue: the rank's own planes
    for b, lo, hi in ((M.Edge, 0, 9), (M.Reflect, 0, 9), (M.Continue, 1, 8)):
        d = M.mapStencil(b, st, arr).desc()
        p = sharding.shard_plan(d, arr.shape, 1, 0)
        assert (p.ghost_lo_src[0], p.ghost_hi_src[0]) == (lo, hi), b
    # fold stencils are centred at 0: the halo is all on the upper side
    d = M.mapStencil(M.Edge, M.sumStencil(M.Sz(3, 2, 2)), arr).desc()
    p = sharding.shard_plan(d, arr.shape, 2, 0)
    assert (p.halo_lo, p.halo_hi) == (0, 2)
    # ... so rank 0 only receives (its upper ghost) and rank 1 only sends (its first planes)
    p1 = sharding.shard_plan(d, arr.shape, 2, 1)
    assert (p.recv_upper_from, p.send_first_to, p.send_last_to, p.recv_lower_from) == (1, -1, -1, -1)
    assert (p1.recv_upper_from, p1.send_first_to, p1.send_last_to, p1.recv_lower_from) == (-1, 0, -1, -1)
    # a star-shaped kernel under Fill 0 is iterated two steps per pass (with or without the finite promise):
    # the slab carries two ghost planes each side although one step reads one
    ks = np.zeros((3, 3, 3))
    ks[1, 1, 1], ks[0, 1, 1], ks[2, 1, 1] = 0.4, 0.1, 0.1
    ds = M.mapStencil(M.Fill(0.0), M.makeConvolutionStencilFromKernel(ks).assumeFinite(), arr).desc()
    ps = sharding.shard_plan(ds, arr.shape, 2, 0)
    assert (ps.halo_lo, ps.halo_hi, ps.step_halo_lo, ps.step_halo_hi) == (2, 2, 1, 1)
    assert list(ps.ghost_lo_src)[:2] == [_abi.GHOST_FILL] * 2 and list(ps.ghost_hi_src)[:2] == [_abi.GHOST_REMOTE] * 2
    assert sharding.shard_plan(M.mapStencil(M.Fill(0.0), M.makeConvolutionStencilFromKernel(ks), arr).desc(), arr.shape, 2, 0).halo_lo == 2
    assert sharding.shard_plan(M.mapStencil(M.Fill(1.0), M.makeConvolutionStencilFromKernel(ks).assumeFinite(), arr).desc(),
                               arr.shape, 2, 0).halo_lo == 1   # non-zero Fill: single steps
    # the halo is the padding: a shifted size-preserving window reads two planes below and none above
    dsh = M.applyStencil(M.Padding((2, 0, 0), (0, 1, 1), M.Fill(0.0)), M.sumStencil(M.Sz(3, 2, 2)), arr).desc()
    psh = sharding.shard_plan(dsh, arr.shape, 2, 1)
    assert (psh.halo_lo, psh.halo_hi, psh.step_halo_lo, psh.step_halo_hi) == (2, 0, 2, 0)
    assert (psh.recv_lower_from, psh.recv_upper_from) == (0, -1)
    # slabs thinner than the halo are refused, as is a shape-changing stencil
    with pytest.raises(M.MassivB200Error):
        sharding.shard_plan(d, arr.shape, 8, 0)
    with pytest.raises(M.SizeMismatchException):
        sharding.shard_plan(M.applyStencil(M.noPadding, st, arr).desc(), arr.shape, 2, 0)
