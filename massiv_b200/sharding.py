"""Slab-sharded iterated stencils: one process per GPU, slabs along the outermost dimension.

The arithmetic of who owns what (``mb_shard_plan``) lives in the C library and needs no GPU.  Two
drivers use it:

* :func:`iterate_sharded` — the product path: rendezvous over ``torch.distributed`` (only to hand
  the NCCL unique id around), then ``mb_iterate_sharded_dev`` which runs the whole loop natively
  (boundary planes -> halo push: copy-engine peer copies into the neighbours' IPC-mapped staging areas, or
  ncclSend/ncclRecv on a second stream when peer memory is unavailable -> interior planes -> join).
* :func:`iterate_sharded_hostcheck` — the same choreography spelled out in Python over any
  ``torch.distributed`` backend (gloo on CPU) with a caller-supplied single-slab compute function.
  It exists so the multi-rank host logic (partition, ghost sources at the physical border, exchange
  order) is testable without GPUs; it is not a compute path of its own.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Sequence

import numpy as np

from . import _abi
from .stencil import Border, DeviceArray, Device, Stencil, mapStencil


def shard_plan(desc: _abi.StencilDesc, global_dims: Sequence[int], n_ranks: int, rank: int) -> _abi.ShardPlan:
    sp = _abi.ShardPlan()
    _abi.raise_for(_abi.lib().mb_shard_plan_make(C.byref(desc), _abi.dims_array(global_dims), n_ranks, rank, C.byref(sp)))
    return sp


def local_valid_desc(desc: _abi.StencilDesc) -> _abi.StencilDesc:
    """the slab problem: the same stencil with NO padding along dim 0 (ghost planes supply the halo)"""
    ld = _abi.StencilDesc.from_buffer_copy(bytes(desc))
    ld._keep = getattr(desc, "_keep", None)
    ld.pad_lo[0] = 0
    ld.pad_hi[0] = 0
    return ld


def init_comm(dev: Device, dist=None) -> None:
    """create the library's NCCL communicator; torch.distributed only carries the unique id"""
    L = dev._L
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        dev.check(L.mb_shard_init(dev.handle, 1, 0, None))
        return
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    idbuf = (C.c_uint8 * _abi.NCCL_ID_BYTES)()
    if rank == 0:
        _abi.raise_for(L.mb_nccl_unique_id(idbuf))
    t = torch.tensor(list(idbuf), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, 0)
    for i, v in enumerate(t.cpu().tolist()):
        idbuf[i] = v
    dev.check(L.mb_shard_init(dev.handle, world, rank, idbuf))


def iterate_sharded(n: int, border: Border, stencil: Stencil, local: DeviceArray, global_dims: Sequence[int],
                    halo_buffers=None):
    """n applications of ``mapStencil border stencil`` to the global array whose slab `local`
    (own planes only) this rank holds.  Returns a DeviceArray view of the rank's own result planes.
    `init_comm` must have been called on the device."""
    dev = local.dev
    L = dev._L
    from .stencil import Padding, applyStencil
    dw = applyStencil(border, stencil._resolved(local.ndim), local) if isinstance(border, Padding) else mapStencil(border, stencil, local)
    desc = dw.desc()
    lo, hi = C.c_int64(), C.c_int64()
    dev.check(L.mb_shard_halo(C.byref(desc), C.byref(lo), C.byref(hi)))
    planes = local.shape[0] + lo.value + hi.value
    plane_elems = int(np.prod(local.shape[1:], dtype=np.int64)) if local.ndim > 1 else 1
    shape = (planes,) + tuple(local.shape[1:])
    if halo_buffers is None:
        a, b = dev.alloc(shape, local.dtype, local.elem), dev.alloc(shape, local.dtype, local.elem)
    else:
        a, b = halo_buffers
    own_off = lo.value * plane_elems * local.dtype.itemsize
    # place the slab after the lower ghost planes (device-to-device through the identity stencil)
    from .stencil import Edge, idStencil
    idd = mapStencil(Edge, idStencil(local.ndim), local).desc()
    dev.check(L.mb_run_dev(dev.handle, C.byref(idd), _abi.dims_array(local.shape), C.c_void_p(local.ptr),
                           C.c_void_p(a.ptr + own_off)))
    res = C.c_void_p()
    dev.check(L.mb_iterate_sharded_dev(dev.handle, C.byref(desc), _abi.dims_array(global_dims),
                                       _abi.dims_array(local.shape), C.c_void_p(a.ptr), C.c_void_p(b.ptr), n, C.byref(res)))
    dev.sync()
    keep = a if res.value == a.ptr else b
    out = dev.wrap(keep.ptr + own_off, local.shape, local.dtype, local.elem)
    out._owner = (a, b)  # keep both buffers alive as long as the view is
    return out


def iterate_sharded_hostcheck(n: int, desc: _abi.StencilDesc, local: np.ndarray, global_dims: Sequence[int], dist,
                              compute: Callable[[_abi.StencilDesc, np.ndarray], np.ndarray], fill_value=0) -> np.ndarray:
    """Python spelling of mb_iterate_sharded_dev's choreography (see module docstring).
    `compute(local_desc, slab_with_ghosts)` must return the slab's own planes after one step."""
    import torch
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    sp = shard_plan(desc, global_dims, world, rank)
    assert local.shape[0] == sp.local_planes
    lo, hi, nl = sp.halo_lo, sp.halo_hi, sp.local_planes
    ld = local_valid_desc(desc)
    buf = np.empty((lo + nl + hi,) + local.shape[1:], local.dtype)
    buf[lo:lo + nl] = local

    def refresh_ghosts(b):
        for j in range(lo):
            s = sp.ghost_lo_src[j]
            if s >= 0:
                b[j] = b[lo + s]
            elif s == _abi.GHOST_FILL:
                b[j] = fill_value
        for j in range(hi):
            s = sp.ghost_hi_src[j]
            if s >= 0:
                b[lo + nl + j] = b[lo + s]
            elif s == _abi.GHOST_FILL:
                b[lo + nl + j] = fill_value
        if max(sp.recv_lower_from, sp.recv_upper_from, sp.send_first_to, sp.send_last_to) < 0:
            return
        ops, recv = [], []
        # same posting order as the native path: sends first-then-last, receives upper-then-lower
        if sp.send_first_to >= 0:
            ops.append(dist.P2POp(dist.isend, torch.from_numpy(np.ascontiguousarray(b[lo:lo + hi])), sp.send_first_to))
        if sp.send_last_to >= 0:
            ops.append(dist.P2POp(dist.isend, torch.from_numpy(np.ascontiguousarray(b[nl:nl + lo])), sp.send_last_to))
        if sp.recv_upper_from >= 0:
            t = torch.empty((hi,) + local.shape[1:], dtype=torch.from_numpy(b[:0]).dtype)
            ops.append(dist.P2POp(dist.irecv, t, sp.recv_upper_from))
            recv.append((slice(lo + nl, lo + nl + hi), t))
        if sp.recv_lower_from >= 0:
            t = torch.empty((lo,) + local.shape[1:], dtype=torch.from_numpy(b[:0]).dtype)
            ops.append(dist.P2POp(dist.irecv, t, sp.recv_lower_from))
            recv.append((slice(0, lo), t))
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        for sl, t in recv:
            b[sl] = t.numpy()

    # the buffers may carry more ghost planes than one step reads (two steps' worth for temporally
    # blocked stencils): a single step sees the view that skips the outermost ones
    rlo, rhi = sp.step_halo_lo, sp.step_halo_hi
    for _ in range(n):
        refresh_ghosts(buf)
        buf[lo:lo + nl] = compute(ld, np.ascontiguousarray(buf[lo - rlo:lo + nl + rhi]))
    return buf[lo:lo + nl].copy()
