// Slab-sharded iteration across the GPUs of one box (one process per GPU).
//
// The reference has no distributed story (single process); this is the analogue of its
// `iterateUntil` loop (Array/Manifest/Internal.hs:331-397) for arrays too large or too slow for one
// GPU.  The global array is cut along the OUTERMOST dimension (contiguous in row-major, so a halo
// plane is one contiguous block).  Per step every rank
//   1. computes the output planes its neighbours need (the first halo_hi and last halo_lo planes),
//   2. hands them to rank-1 / rank+1 with ncclSend/ncclRecv on a second stream,
//   3. computes the interior planes on the main stream while the exchange is in flight,
//   4. joins the two streams and swaps buffers.
// The physical border of the global array is realised in the ghost planes: Fill -> constant planes,
// Wrap -> ring neighbours, Edge/Reflect/Continue -> copies of the rank's own planes (handleBorderIndex
// is per-dimension, Core/Index/Internal.hs:653-658, so a resolved ghost plane is an ordinary plane).
// NCCL is resolved with dlopen at first use so that single-GPU users never need it.
#include <dlfcn.h>

#include <cstdlib>
#include <nccl.h>

#include "mb_internal.h"

namespace mb {

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

static NcclApi &nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // reuse an already loaded NCCL (e.g. the one torch bundles), else the system library
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.lib = h;
#define MB_SYM(field, name) api.field = (decltype(api.field))dlsym(h, name)
    MB_SYM(GetUniqueId, "ncclGetUniqueId");
    MB_SYM(CommInitRank, "ncclCommInitRank");
    MB_SYM(CommDestroy, "ncclCommDestroy");
    MB_SYM(Send, "ncclSend");
    MB_SYM(Recv, "ncclRecv");
    MB_SYM(GroupStart, "ncclGroupStart");
    MB_SYM(GroupEnd, "ncclGroupEnd");
    MB_SYM(GetErrorString, "ncclGetErrorString");
#undef MB_SYM
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Send && api.Recv && api.GroupStart &&
             api.GroupEnd;
  });
  return api;
}

struct ShardState {
  ncclComm_t comm = nullptr;
  int n = 1, rank = 0;
};

void shard_destroy(Ctx *c) {
  if (!c->shard) return;
  if (c->shard->comm && nccl().ok) nccl().CommDestroy(c->shard->comm);
  delete c->shard;
  c->shard = nullptr;
}

static int check_nccl(Ctx *c, ncclResult_t r, const char *what) {
  if (r == ncclSuccess) return MB_OK;
  const char *s = nccl().GetErrorString ? nccl().GetErrorString(r) : "?";
  return fail(c, MB_ERR_NCCL, std::string(what) + ": " + s);
}
#define MB_NCCL(ctx, expr)                              \
  do {                                                  \
    int _st = check_nccl((ctx), (expr), #expr);         \
    if (_st) return _st;                                \
  } while (0)

static inline int64_t floor_mod(int64_t a, int64_t k) {
  int64_t m = a % k;
  return (m != 0 && ((m < 0) != (k < 0))) ? m + k : m;
}

// Where do rank `rank`'s ghost planes come from?  Ghost plane j of the lower halo is global plane
// first - halo_lo + j; ghost plane j of the upper halo is global plane first + local + j.
static int ghost_sources(const mb_stencil_desc *d, const Plan &p, int64_t N, int n_ranks, int rank, mb_shard_plan *out,
                         std::string *err) {
  auto bad = [&](int st, const char *m) { if (err) *err = m; return st; };
  std::memset(out, 0, sizeof(*out));
  out->first_plane = (int64_t)rank * N / n_ranks;
  out->local_planes = (int64_t)(rank + 1) * N / n_ranks - out->first_plane;
  out->step_halo_lo = p.center[0];
  out->step_halo_hi = p.geom[0] - 1 - p.center[0];
  if (out->step_halo_hi < 0) out->step_halo_hi = 0;
  // temporally blocked stencils advance two steps per exchange and therefore keep two steps' worth
  // of ghost planes
  const int64_t depth = tb2_desc_eligible(d) ? 2 : 1;
  out->halo_lo = out->step_halo_lo * depth;
  out->halo_hi = out->step_halo_hi * depth;
  if (out->halo_lo > MB_MAX_HALO || out->halo_hi > MB_MAX_HALO) return bad(MB_ERR_UNSUPPORTED, "halo deeper than MB_MAX_HALO planes");
  const int64_t need = out->halo_lo > out->halo_hi ? out->halo_lo : out->halo_hi;
  if (N < n_ranks || out->local_planes < need || out->local_planes < 1)
    return bad(MB_ERR_UNSUPPORTED, "slab thinner than the stencil halo: use fewer ranks");
  out->recv_lower_from = out->recv_upper_from = out->send_first_to = out->send_last_to = -1;
  for (int side = 0; side < 2; ++side) {
    const int64_t h = side == 0 ? out->halo_lo : out->halo_hi;
    int64_t *src = side == 0 ? out->ghost_lo_src : out->ghost_hi_src;
    int32_t *nbr = side == 0 ? &out->recv_lower_from : &out->recv_upper_from;
    for (int64_t j = 0; j < h; ++j) {
      int64_t g = side == 0 ? out->first_plane - h + j : out->first_plane + out->local_planes + j;
      src[j] = MB_GHOST_REMOTE;
      if (g < 0 || g >= N) {  // beyond the global array: the stencil's Border decides
        int64_t rep; int32_t is_fill;
        int bs = border_index(d->border, N, g, &rep, &is_fill);
        if (bs) return bad(bs, "border resolution failed");
        if (is_fill) { src[j] = MB_GHOST_FILL; continue; }
        g = rep;
      }
      if (g >= out->first_plane && g < out->first_plane + out->local_planes) {
        src[j] = g - out->first_plane;  // one of my own planes (physical border or 1-rank wrap)
        continue;
      }
      // must be owned by the adjacent rank (ring order for Wrap)
      const int want = side == 0 ? (rank + n_ranks - 1) % n_ranks : (rank + 1) % n_ranks;
      const int64_t wf = (int64_t)want * N / n_ranks, wl = (int64_t)(want + 1) * N / n_ranks;
      const int64_t expect = side == 0 ? wl - h + j : wf + j;
      if (g != expect) return bad(MB_ERR_UNSUPPORTED, "ghost plane is not on an adjacent rank: use fewer ranks");
      *nbr = want;
    }
  }
  return MB_OK;
}

// The slab plan (host only): my ghost sources, plus whom my boundary planes must be sent to (the
// adjacent ranks whose ghost sources point at me).
int shard_plan(const mb_stencil_desc *d, const int64_t *global_dims, int n_ranks, int rank, mb_shard_plan *out,
               std::string *err) {
  auto bad = [&](int st, const char *m) { if (err) *err = m; return st; };
  if (!d || !global_dims || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks)
    return bad(MB_ERR_INVALID_DESC, "bad shard arguments");
  Plan p;
  int st = plan_geometry(d, global_dims, &p, err);
  if (st) return st;
  if (!p.same_shape() || !p.unit_stride())
    return bad(MB_ERR_SIZE_MISMATCH, "sharded iteration needs a size-preserving (mapStencil) stencil");
  const int64_t N = global_dims[0];
  if ((st = ghost_sources(d, p, N, n_ranks, rank, out, err))) return st;
  if (n_ranks > 1) {
    mb_shard_plan q;
    const int lower = (rank + n_ranks - 1) % n_ranks, upper = (rank + 1) % n_ranks;
    if ((st = ghost_sources(d, p, N, n_ranks, lower, &q, err))) return st;
    if (q.recv_upper_from == rank) out->send_first_to = lower;
    if ((st = ghost_sources(d, p, N, n_ranks, upper, &q, err))) return st;
    if (q.recv_lower_from == rank) out->send_last_to = upper;
  }
  return MB_OK;
}

// make `buf`'s ghost planes valid: local copies / fills on `stream`, neighbour planes by NCCL on comm_stream
struct SlabGeom {
  size_t plane_bytes;
  int64_t local, lo, hi;
};

static int fill_local_ghosts(Ctx *c, const mb_shard_plan &sp, const SlabGeom &g, char *buf, const DevPlan &fillplan,
                             bool fills_too, cudaStream_t stream) {
  (void)fillplan;
  for (int side = 0; side < 2; ++side) {
    const int64_t h = side == 0 ? g.lo : g.hi;
    const int64_t *src = side == 0 ? sp.ghost_lo_src : sp.ghost_hi_src;
    for (int64_t j = 0; j < h; ++j) {
      char *dst = buf + (size_t)(side == 0 ? j : g.lo + g.local + j) * g.plane_bytes;
      if (src[j] >= 0) {
        MB_CUDA(c, cudaMemcpyAsync(dst, buf + (size_t)(g.lo + src[j]) * g.plane_bytes, g.plane_bytes,
                                   cudaMemcpyDeviceToDevice, stream));
      } else if (src[j] == MB_GHOST_FILL && fills_too) {
        // constant planes never change: written once per buffer by the caller (fill_const_plane)
      }
    }
  }
  return MB_OK;
}

__global__ void fill_plane_kernel(unsigned char *dst, size_t n_elems, int esize, unsigned long long bits) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_elems; i += (size_t)gridDim.x * blockDim.x) {
    for (int b = 0; b < esize; ++b) dst[i * esize + b] = (unsigned char)(bits >> (8 * b));
  }
}

static int exchange(Ctx *c, const mb_shard_plan &sp, const SlabGeom &g, char *buf, cudaStream_t stream) {
  ShardState *S = c->shard;
  if (sp.recv_lower_from < 0 && sp.recv_upper_from < 0 && sp.send_first_to < 0 && sp.send_last_to < 0) return MB_OK;
  NcclApi &api = nccl();
  char *first = buf + (size_t)g.lo * g.plane_bytes;                       // my first planes -> lower rank's upper ghost
  char *last = buf + (size_t)(g.lo + g.local - g.lo) * g.plane_bytes;     // my last halo_lo planes -> upper rank's lower ghost
  char *ghost_lo = buf;
  char *ghost_hi = buf + (size_t)(g.lo + g.local) * g.plane_bytes;
  MB_NCCL(c, api.GroupStart());
  // order matters when both neighbours are the same rank (2-rank ring): sends lower-then-upper,
  // receives upper-then-lower, so the peer's first send lands in my upper ghost
  if (sp.send_first_to >= 0) MB_NCCL(c, api.Send(first, (size_t)g.hi * g.plane_bytes, ncclUint8, sp.send_first_to, S->comm, stream));
  if (sp.send_last_to >= 0) MB_NCCL(c, api.Send(last, (size_t)g.lo * g.plane_bytes, ncclUint8, sp.send_last_to, S->comm, stream));
  if (sp.recv_upper_from >= 0) MB_NCCL(c, api.Recv(ghost_hi, (size_t)g.hi * g.plane_bytes, ncclUint8, sp.recv_upper_from, S->comm, stream));
  if (sp.recv_lower_from >= 0) MB_NCCL(c, api.Recv(ghost_lo, (size_t)g.lo * g.plane_bytes, ncclUint8, sp.recv_lower_from, S->comm, stream));
  MB_NCCL(c, api.GroupEnd());
  return MB_OK;
}

}  // namespace mb

using namespace mb;

extern "C" {

int mb_shard_plan_make(const mb_stencil_desc *desc, const int64_t *global_dims, int n_ranks, int rank,
                       mb_shard_plan *out) {
  return shard_plan(desc, global_dims, n_ranks, rank, out, nullptr);
}

int mb_shard_halo(const mb_stencil_desc *desc, int64_t *halo_lo, int64_t *halo_hi) {
  if (!desc || !halo_lo || !halo_hi) return MB_ERR_INVALID_DESC;
  int64_t dims[MB_MAX_RANK];
  for (int i = 0; i < MB_MAX_RANK; ++i) dims[i] = 1 << 20;
  mb_shard_plan sp;
  int st = shard_plan(desc, dims, 1, 0, &sp, nullptr);
  if (st) return st;
  *halo_lo = sp.halo_lo;
  *halo_hi = sp.halo_hi;
  return MB_OK;
}

int mb_nccl_unique_id(uint8_t *id) {
  if (!id) return MB_ERR_INVALID_DESC;
  NcclApi &api = nccl();
  if (!api.ok) return MB_ERR_NCCL;
  static_assert(sizeof(ncclUniqueId) <= MB_NCCL_ID_BYTES, "ncclUniqueId larger than MB_NCCL_ID_BYTES");
  ncclUniqueId u;
  if (api.GetUniqueId(&u) != ncclSuccess) return MB_ERR_NCCL;
  std::memset(id, 0, MB_NCCL_ID_BYTES);
  std::memcpy(id, &u, sizeof(u));
  return MB_OK;
}

int mb_shard_init(mb_ctx *h, int n_ranks, int rank, const uint8_t *id) {
  if (!h) return MB_ERR_INVALID_DESC;
  Ctx *c = &h->c;
  std::lock_guard<std::mutex> lk(c->mu);
  MB_CUDA(c, cudaSetDevice(c->device));
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(c, MB_ERR_INVALID_DESC, "bad rank / world size");
  shard_destroy(c);
  c->shard = new ShardState();
  c->shard->n = n_ranks;
  c->shard->rank = rank;
  if (n_ranks == 1) return MB_OK;  // nothing to talk to: ghosts are local copies
  if (!id) return fail(c, MB_ERR_INVALID_DESC, "missing NCCL unique id");
  NcclApi &api = nccl();
  if (!api.ok) return fail(c, MB_ERR_NCCL, "libnccl.so.2 could not be loaded");
  ncclUniqueId u;
  std::memcpy(&u, id, sizeof(u));
  MB_NCCL(c, api.CommInitRank(&c->shard->comm, n_ranks, u, rank));
  return MB_OK;
}

static int iterate_sharded_locked(Ctx *c, const mb_stencil_desc *desc, const int64_t *global_dims,
                                  const int64_t *local_dims, void *dev_a, void *dev_b, int64_t n_steps, void **result);

int mb_iterate_sharded_dev(mb_ctx *h, const mb_stencil_desc *desc, const int64_t *global_dims,
                           const int64_t *local_dims, void *dev_a, void *dev_b, int64_t n_steps, void **result) {
  if (!h) return MB_ERR_INVALID_DESC;
  Ctx *c = &h->c;
  std::lock_guard<std::mutex> lk(c->mu);
  MB_CUDA(c, cudaSetDevice(c->device));
  return iterate_sharded_locked(c, desc, global_dims, local_dims, dev_a, dev_b, n_steps, result);
}

/* host -> host: this rank's slab (own planes only) in, own planes out */
int mb_iterate_sharded(mb_ctx *h, const mb_stencil_desc *desc, const int64_t *global_dims, const int64_t *local_dims,
                       const void *host_in, void *host_out, int64_t n_steps) {
  if (!h) return MB_ERR_INVALID_DESC;
  Ctx *c = &h->c;
  std::lock_guard<std::mutex> lk(c->mu);
  MB_CUDA(c, cudaSetDevice(c->device));
  if (!desc || !global_dims || !local_dims) return fail(c, MB_ERR_INVALID_DESC, "null argument");
  if (!c->shard) return fail(c, MB_ERR_INVALID_DESC, "mb_shard_init has not been called");
  mb_shard_plan sp;
  std::string err;
  int st = shard_plan(desc, global_dims, c->shard->n, c->shard->rank, &sp, &err);
  if (st) return fail(c, st, err);
  size_t plane_bytes = (size_t)mb_elem_size(desc->elem);
  for (int i = 1; i < desc->rank; ++i) plane_bytes *= (size_t)local_dims[i];
  const size_t own = plane_bytes * (size_t)local_dims[0];
  const size_t total = plane_bytes * (size_t)(sp.halo_lo + local_dims[0] + sp.halo_hi);
  void *a = nullptr, *b = nullptr, *res = nullptr;
  if ((st = ensure_scratch(c, 0, total ? total : 1, &a))) return st;
  if ((st = ensure_scratch(c, 1, total ? total : 1, &b))) return st;
  if (own) MB_CUDA(c, cudaMemcpyAsync((char *)a + sp.halo_lo * plane_bytes, host_in, own, cudaMemcpyHostToDevice, c->stream));
  if ((st = iterate_sharded_locked(c, desc, global_dims, local_dims, a, b, n_steps, &res))) return st;
  if (own) MB_CUDA(c, cudaMemcpyAsync(host_out, (char *)res + sp.halo_lo * plane_bytes, own, cudaMemcpyDeviceToHost, c->stream));
  MB_CUDA(c, cudaStreamSynchronize(c->stream));
  return MB_OK;
}

}  // extern "C"

static int iterate_sharded_locked(Ctx *c, const mb_stencil_desc *desc, const int64_t *global_dims,
                                  const int64_t *local_dims, void *dev_a, void *dev_b, int64_t n_steps, void **result) {
  if (!desc || !global_dims || !local_dims || !result) return fail(c, MB_ERR_INVALID_DESC, "null argument");
  if (!c->shard) return fail(c, MB_ERR_INVALID_DESC, "mb_shard_init has not been called");
  if (n_steps < 1) return fail(c, MB_ERR_INVALID_DESC, "iterateUntil applies the step at least once");
  ShardState *S = c->shard;
  mb_shard_plan sp;
  std::string err;
  int st = shard_plan(desc, global_dims, S->n, S->rank, &sp, &err);
  if (st) return fail(c, st, err);
  if (local_dims[0] != sp.local_planes) return fail(c, MB_ERR_SIZE_MISMATCH, "local_dims[0] does not match the slab plan");
  for (int i = 1; i < desc->rank; ++i)
    if (local_dims[i] != global_dims[i]) return fail(c, MB_ERR_SIZE_MISMATCH, "inner dimensions are not sharded");

  // the local problem: the same stencil applied with NO padding along dim 0 to the slab + ghosts.
  // The buffers may carry more ghost planes than one step needs (extra_lo / extra_hi, temporal
  // blocking): a single step then reads a view that skips the outermost ones.
  mb_stencil_desc ld = *desc;
  ld.pad_lo[0] = 0;
  ld.pad_hi[0] = 0;
  const int64_t extra_lo = sp.halo_lo - sp.step_halo_lo, extra_hi = sp.halo_hi - sp.step_halo_hi;
  int64_t in_dims[MB_MAX_RANK];
  for (int i = 0; i < desc->rank; ++i) in_dims[i] = local_dims[i];
  in_dims[0] = sp.step_halo_lo + sp.local_planes + sp.step_halo_hi;
  std::shared_ptr<DevPlan> dp, dp2;
  if ((st = get_dev_plan(c, &ld, in_dims, &dp))) return st;
  if (dp->p.out_dims[0] != sp.local_planes) return fail(c, MB_ERR_SIZE_MISMATCH, "internal: slab geometry");
  bool use_tb2 = false;
  if (extra_lo > 0 && extra_lo == sp.step_halo_lo && extra_hi == sp.step_halo_hi && sp.halo_lo == 2 && sp.halo_hi == 2) {
    in_dims[0] = sp.halo_lo + sp.local_planes + sp.halo_hi;
    if ((st = get_dev_plan(c, &ld, in_dims, &dp2))) return st;
    use_tb2 = vol3d_tb2_applicable(c, dp2->p);
  }

  SlabGeom g;
  g.plane_bytes = (size_t)dp->p.esize;
  for (int i = 1; i < desc->rank; ++i) g.plane_bytes *= (size_t)local_dims[i];
  g.local = sp.local_planes; g.lo = sp.halo_lo; g.hi = sp.halo_hi;
  char *bufs[2] = {(char *)dev_a, (char *)dev_b};

  // constant ghost planes (Fill at a physical border) are written once into both buffers
  unsigned long long fill_bits = 0;
  std::memcpy(&fill_bits, desc->fill, 8);
  const size_t plane_elems = g.plane_bytes / dp->p.esize;
  for (int b = 0; b < 2; ++b)
    for (int side = 0; side < 2; ++side) {
      const int64_t hh = side == 0 ? g.lo : g.hi;
      const int64_t *src = side == 0 ? sp.ghost_lo_src : sp.ghost_hi_src;
      for (int64_t j = 0; j < hh; ++j)
        if (src[j] == MB_GHOST_FILL) {
          unsigned char *dst = (unsigned char *)bufs[b] + (size_t)(side == 0 ? j : g.lo + g.local + j) * g.plane_bytes;
          fill_plane_kernel<<<c->num_sms * 4, 256, 0, c->stream>>>(dst, plane_elems, dp->p.esize, fill_bits);
          c->launches++;
        }
    }
  MB_CUDA(c, cudaGetLastError());

  // ghosts of the initial state
  if ((st = fill_local_ghosts(c, sp, g, bufs[0], *dp, false, c->stream))) return st;
  MB_CUDA(c, cudaEventRecord(c->ev_a, c->stream));
  MB_CUDA(c, cudaStreamWaitEvent(c->comm_stream, c->ev_a, 0));
  if ((st = exchange(c, sp, g, bufs[0], c->comm_stream))) return st;
  MB_CUDA(c, cudaEventRecord(c->ev_b, c->comm_stream));
  MB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_b, 0));

  const bool talks = sp.recv_lower_from >= 0 || sp.recv_upper_from >= 0 || sp.send_first_to >= 0 || sp.send_last_to >= 0 ||
                     getenv("MASSIV_B200_FORCE_SPLIT") != nullptr;  // experiment knob: peel boundary planes even when alone
  int cur = 0;
  for (int64_t s = 0; s < n_steps;) {
    const bool two = use_tb2 && n_steps - s >= 2;  // a fused pair of steps, or one step
    const int64_t adv = two ? 2 : 1;
    char *in = bufs[cur], *out_base = bufs[1 - cur];
    char *out = out_base + (size_t)g.lo * g.plane_bytes;  // the rank's own planes start after the lower ghost
    const bool last = s + adv == n_steps;
    auto launch = [&](const OuterRange &r) -> int {
      if (r.o1 <= r.o0) return MB_OK;
      if (two) return launch_vol3d_tb2(c, dp2->p, in, out, &r, (int)g.lo, sp.first_plane - g.lo, global_dims[0]);
      return launch_plan(c, *dp, in + (size_t)extra_lo * g.plane_bytes, out, &r);
    };
    // 1. the planes the neighbours are waiting for (a full ghost depth at either end) first
    int64_t top_end = g.hi < g.local ? g.hi : g.local;                        // result planes [0, top_end)
    int64_t bot_begin = g.local - g.lo > top_end ? g.local - g.lo : top_end;  // result planes [bot_begin, local)
    if (last || !talks) { top_end = 0; bot_begin = g.local; }  // nobody waits: one launch
    const OuterRange r1{0, top_end}, r2{bot_begin, g.local}, r3{top_end, bot_begin};
    if ((st = launch(r1))) return st;
    if ((st = launch(r2))) return st;
    if (!last) {
      MB_CUDA(c, cudaEventRecord(c->ev_a, c->stream));
      MB_CUDA(c, cudaStreamWaitEvent(c->comm_stream, c->ev_a, 0));
      // 2. neighbour exchange of the new boundary planes on the second stream, concurrent with
      if ((st = exchange(c, sp, g, out_base, c->comm_stream))) return st;
      MB_CUDA(c, cudaEventRecord(c->ev_b, c->comm_stream));
    }
    // 3. the interior
    if ((st = launch(r3))) return st;
    if (!last) {
      // physical-border ghosts that are copies of own planes (Edge / Reflect / Continue / 1-rank Wrap)
      if ((st = fill_local_ghosts(c, sp, g, out_base, *dp, false, c->stream))) return st;
      // 4. join
      MB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_b, 0));
    }
    cur = 1 - cur;
    s += adv;
  }
  *result = bufs[cur];
  return MB_OK;
}
