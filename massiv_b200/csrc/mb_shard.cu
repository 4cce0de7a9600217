// Slab-sharded iteration across the GPUs of one box (one process per GPU).
//
// The reference has no distributed story (single process); this is the analogue of its
// `iterateUntil` loop (Array/Manifest/Internal.hs:331-397) for arrays too large or too slow for one
// GPU.  The global array is cut along the OUTERMOST dimension (contiguous in row-major, so a halo
// plane is one contiguous block).  Per round (one step, or two fused steps for star-shaped 3-D kernels
// under Fill 0) every rank
//   1. takes delivery of its neighbours' planes for this round,
//   2. computes the output planes its neighbours need (the first halo_hi and last halo_lo planes),
//   3. ships them — peer copies over NVLink into the neighbours' staging areas with counter hand-shakes
//      in device memory (below), or ncclSend/ncclRecv on a second stream as the fallback,
//   4. computes the interior planes while the transfer is in flight, and swaps buffers.
// (Promise-free fused rounds compute the slab in one launch, check it, and ship afterwards.)
// The physical border of the global array is realised in the ghost planes: Fill -> constant planes,
// Wrap -> ring neighbours, Edge/Reflect/Continue -> copies of the rank's own planes (handleBorderIndex
// is per-dimension, Core/Index/Internal.hs:653-658, so a resolved ghost plane is an ordinary plane).
// NCCL is resolved with dlopen at first use so that single-GPU users never need it.
#include <cuda.h>
#include <dlfcn.h>

#include <cstdlib>
#include <map>
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

// ---- peer-memory halo exchange ---------------------------------------------------------------------
// NCCL's send/recv kernel needs SM slots that the persistent interior kernel occupies, so the exchange
// only ran once that kernel had drained and cost 1.8 ms per step at 4 GPUs.  The halo planes are
// therefore moved by the copy engines over NVLink into a small staging area that every rank owns for
// the life of its communicator and that its neighbours map with CUDA IPC (the caller's slab buffers are
// never exported: their lifetime is the caller's business).  Two ranks hand-shake through 32-bit
// counters in device memory, numbered by round:
//   delivered[side]  written by the neighbour after its copy landed: "your ghosts for round r are staged"
//   credit[side]     written by the neighbour after it copied them out: "that staging slot is free again"
// Waiting is a stream memory operation (cuStreamWaitValue32), so no SM is involved anywhere; the
// receiver moves staged planes into its ghost planes with a local copy (4 planes per round).
// NCCL remains the rendezvous (IPC handles travel over it) and the fallback.
enum { F_DELIV_LO = 0, F_DELIV_HI = 1, F_CREDIT_LO = 2, F_CREDIT_HI = 3, F_COUNT = 16 };
static const size_t STAGE_HEADER = 256;  // the staging area starts with a magic word, plane slots follow

struct IpcInfo {  // what a rank tells its neighbours at the start of an iterate call
  cudaIpcMemHandle_t stage;
  cudaIpcMemHandle_t flags;
  unsigned long long stage_off, flags_off;  // of the pointer inside the exported allocation
  unsigned long long stage_bytes;
  int ok;
  int pad;
};

struct PeerView {
  char *stage = nullptr;
  unsigned *flags = nullptr;
};

typedef CUresult (*WaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*GetAddressRangeFn)(CUdeviceptr *, size_t *, CUdeviceptr);

struct ShardState {
  ncclComm_t comm = nullptr;
  int n = 1, rank = 0;
  // peer-memory path
  unsigned *flags = nullptr;        // F_COUNT counters (+ a magic word), set to `rounds` when a call starts
  char *xstage = nullptr;           // staging area the neighbours write my ghost planes into
  size_t xstage_bytes = 0;
  std::vector<void *> retired;      // outgrown staging areas: freed with the communicator (a peer may still map them)
  IpcInfo *stage = nullptr;         // device scratch for the handle exchange: [0] mine, [1] from lower, [2] from upper
  unsigned rounds = 0;              // peer-memory rounds completed on this communicator (identical on every rank)
  std::map<std::string, void *> opened;  // IPC handle bytes -> mapped base
  WaitValue32Fn wait_value = nullptr;
  GetAddressRangeFn addr_range = nullptr;
  bool p2p_tried = false, p2p_ok = false;
  const char *transport = "none";
};

void shard_destroy(Ctx *c) {
  if (!c->shard) return;
  for (auto &kv : c->shard->opened) cudaIpcCloseMemHandle(kv.second);
  if (c->shard->flags) cudaFree(c->shard->flags);
  if (c->shard->xstage) cudaFree(c->shard->xstage);
  for (void *p : c->shard->retired) cudaFree(p);
  if (c->shard->stage) cudaFree(c->shard->stage);
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
  // the planes ONE application reads below / above the slab.  Output plane o evaluates the stencil at source
  // plane o + (centre - pad_lo) (Array/Stencil.hs:200), so its window covers o - pad_lo .. o + pad_hi for a
  // size-preserving descriptor (pad_lo + pad_hi = size - 1): the PADDING is the halo.  For mapStencil
  // (pad_lo = centre) this is the familiar centre / size-1-centre.
  out->step_halo_lo = p.pad_lo[0];
  out->step_halo_hi = p.pad_hi[0];
  // temporally blocked stencils advance two steps per exchange and therefore keep two steps' worth
  // of ghost planes
  const int64_t depth = (tb2_desc_eligible(d) && p.pad_lo[0] == 1 && p.pad_hi[0] == 1) ? 2 : 1;
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

// ---- peer-memory path: set-up -------------------------------------------------------------------------

__global__ void set_flag_kernel(unsigned *flag, unsigned value) {
  *reinterpret_cast<volatile unsigned *>(flag) = value;
  __threadfence_system();
}
__global__ void spin_flag_kernel(const unsigned *flag, unsigned value) {  // only when stream memory ops are missing
  while ((int)(*reinterpret_cast<const volatile unsigned *>(flag) - value) < 0) __nanosleep(200);
  __threadfence_system();
}

struct P2P {
  bool on = false;
  PeerView lower, upper;  // the neighbours' staging areas / counters as mapped into this process
};

static void p2p_note(const ShardState *S, const char *what) {
  static const bool dbg = getenv("MASSIV_B200_DEBUG_P2P") != nullptr;
  if (dbg) fprintf(stderr, "[massiv_b200 rank %d] peer-memory halo exchange: %s\n", S->rank, what);
}

static void *open_ipc(ShardState *S, const cudaIpcMemHandle_t &h, std::vector<std::string> *used) {
  const std::string key((const char *)&h, sizeof(h));
  used->push_back(key);
  auto it = S->opened.find(key);
  if (it != S->opened.end()) return it->second;
  void *p = nullptr;
  if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  S->opened[key] = p;
  return p;
}

static unsigned long long p2p_magic(int rank, int which) {
  return 0x6d61737369760000ull + (unsigned long long)rank * 16 + (unsigned long long)which;
}

// where the planes bound for ghost side `side` (0 = lower, 1 = upper) of buffer parity `p` are staged
static inline size_t stage_slot(const SlabGeom &g, int p, int side) {
  return STAGE_HEADER + ((size_t)p * (size_t)(g.lo + g.hi) + (side ? (size_t)g.lo : 0)) * g.plane_bytes;
}

// Make sure my staging area is large enough, swap its IPC handle (and the counter block's) with the
// neighbours over NCCL, map theirs and check the mappings.  Every rank ends up with the same answer (all
// use peer memory, or all use NCCL): a rank that cannot export or map says so, and the verdict is relayed
// along the chain of neighbours.
static int setup_p2p(Ctx *c, const mb_shard_plan &sp, const SlabGeom &g, P2P *out) {
  ShardState *S = c->shard;
  out->on = false;
  if (S->n < 2) return MB_OK;
  NcclApi &api = nccl();
  // A rank that does not want (MASSIV_B200_NO_P2P / option "no_p2p") or cannot use peer memory still takes
  // part in the handle exchange and the consensus below, saying "no": whether peer memory is used is decided
  // from state only one process knows, and a rank that simply left would leave its neighbours waiting.
  const bool unwanted = c->no_p2p || getenv("MASSIV_B200_NO_P2P") != nullptr;
  if (!S->p2p_tried) {
    S->p2p_tried = true;
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fp, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      S->wait_value = (WaitValue32Fn)fp;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      S->addr_range = (GetAddressRangeFn)fp;
    bool ok = S->addr_range != nullptr;
    // allocations of 2 MiB and more are blocks of their own, so their IPC handles are unambiguous
    ok = ok && cudaMalloc(&S->flags, (size_t)2 << 20) == cudaSuccess;
    ok = ok && cudaMemset(S->flags, 0, (size_t)2 << 20) == cudaSuccess;
    S->p2p_ok = ok;
    if (!ok) p2p_note(S, "cannot allocate the counter block");
    cudaGetLastError();
  }
  if (!S->stage) return fail(c, MB_ERR_OOM, "internal: the handle-exchange scratch was not allocated by mb_shard_init");
  // 1. my side: staging area, counters set to the communicator's round number, magic words
  IpcInfo mine;
  std::memset(&mine, 0, sizeof(mine));
  mine.ok = (S->p2p_ok && !unwanted) ? 1 : 0;
  const size_t need = STAGE_HEADER + 2 * (size_t)(g.lo + g.hi) * g.plane_bytes;
  if (mine.ok && S->xstage_bytes < need) {
    if (S->xstage) S->retired.push_back(S->xstage);
    S->xstage = nullptr;
    S->xstage_bytes = 0;
    const size_t bytes = need < ((size_t)2 << 20) ? ((size_t)2 << 20) : need;
    if (cudaMalloc(&S->xstage, bytes) == cudaSuccess) S->xstage_bytes = bytes;
    else { cudaGetLastError(); mine.ok = 0; p2p_note(S, "cannot allocate the staging area"); }
  }
  if (mine.ok) {
    for (int f = 0; f < 4; ++f) set_flag_kernel<<<1, 1, 0, c->stream>>>(&S->flags[f], S->rounds);
    const unsigned long long m0 = p2p_magic(S->rank, 0), m1 = p2p_magic(S->rank, 1);
    MB_CUDA(c, cudaMemcpyAsync(S->xstage, &m0, 8, cudaMemcpyHostToDevice, c->stream));
    MB_CUDA(c, cudaMemcpyAsync(S->flags + F_COUNT, &m1, 8, cudaMemcpyHostToDevice, c->stream));
    CUdeviceptr base = 0;
    size_t size = 0;
    if (S->addr_range(&base, &size, (CUdeviceptr)S->xstage) != CUDA_SUCCESS ||
        cudaIpcGetMemHandle(&mine.stage, (void *)base) != cudaSuccess) { mine.ok = 0; cudaGetLastError(); }
    mine.stage_off = (unsigned long long)((CUdeviceptr)S->xstage - base);
    if (mine.ok && (S->addr_range(&base, &size, (CUdeviceptr)S->flags) != CUDA_SUCCESS ||
                    cudaIpcGetMemHandle(&mine.flags, (void *)base) != cudaSuccess)) { mine.ok = 0; cudaGetLastError(); }
    mine.flags_off = (unsigned long long)((CUdeviceptr)S->flags - base);
    mine.stage_bytes = S->xstage_bytes;
    if (!mine.ok) p2p_note(S, "cannot export device memory (cudaIpcGetMemHandle)");
  }
  // 2. swap with the ring neighbours (ranks at a physical Fill/Edge border still have a chain neighbour)
  const int lower = sp.recv_lower_from >= 0 ? sp.recv_lower_from : sp.send_first_to;
  const int upper = sp.recv_upper_from >= 0 ? sp.recv_upper_from : sp.send_last_to;
  MB_CUDA(c, cudaMemcpyAsync(&S->stage[0], &mine, sizeof(mine), cudaMemcpyHostToDevice, c->stream));
  MB_NCCL(c, api.GroupStart());
  if (lower >= 0) MB_NCCL(c, api.Send(&S->stage[0], sizeof(IpcInfo), ncclUint8, lower, S->comm, c->stream));
  if (upper >= 0) MB_NCCL(c, api.Send(&S->stage[0], sizeof(IpcInfo), ncclUint8, upper, S->comm, c->stream));
  if (upper >= 0) MB_NCCL(c, api.Recv(&S->stage[2], sizeof(IpcInfo), ncclUint8, upper, S->comm, c->stream));
  if (lower >= 0) MB_NCCL(c, api.Recv(&S->stage[1], sizeof(IpcInfo), ncclUint8, lower, S->comm, c->stream));
  MB_NCCL(c, api.GroupEnd());
  IpcInfo got[2];
  std::memset(got, 0, sizeof(got));
  MB_CUDA(c, cudaMemcpyAsync(got, &S->stage[1], 2 * sizeof(IpcInfo), cudaMemcpyDeviceToHost, c->stream));
  MB_CUDA(c, cudaStreamSynchronize(c->stream));
  // 3. map and verify
  int good = mine.ok;
  PeerView lo, up;
  std::vector<std::string> used;
  auto map_peer = [&](const IpcInfo &info, int peer_rank, PeerView *v) {
    if (!info.ok) { good = 0; return; }
    if (info.stage_bytes < need) { good = 0; p2p_note(S, "a neighbour's staging area is too small"); return; }
    char *sb = (char *)open_ipc(S, info.stage, &used);
    char *fb = (char *)open_ipc(S, info.flags, &used);
    if (!sb || !fb) { good = 0; p2p_note(S, "cudaIpcOpenMemHandle failed"); return; }
    v->stage = sb + info.stage_off;
    v->flags = (unsigned *)(fb + info.flags_off);
    unsigned long long seen[2] = {0, 0};
    if (cudaMemcpy(&seen[0], v->stage, 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(&seen[1], v->flags + F_COUNT, 8, cudaMemcpyDeviceToHost) != cudaSuccess) {
      cudaGetLastError(); good = 0; p2p_note(S, "cannot read mapped peer memory"); return;
    }
    if (seen[0] != p2p_magic(peer_rank, 0) || seen[1] != p2p_magic(peer_rank, 1)) { good = 0; p2p_note(S, "mapped peer memory is not the neighbour's"); }
  };
  if (lower >= 0) map_peer(got[0], lower, &lo);
  if (upper >= 0) map_peer(got[1], upper, &up);
  // mappings of staging areas the neighbours have outgrown are released (they keep the memory itself
  // until their communicator goes away, so closing late is safe)
  if (good) {
    for (auto it = S->opened.begin(); it != S->opened.end();) {
      bool keep = false;
      for (const std::string &u : used) if (u == it->first) keep = true;
      if (keep) { ++it; } else { cudaIpcCloseMemHandle(it->second); it = S->opened.erase(it); }
    }
    cudaGetLastError();
  }
  // 4. consensus: n_ranks-1 rounds of neighbour exchanges of one word reach every rank of a chain or ring
  int *flagdev = reinterpret_cast<int *>(&S->stage[0]);
  for (int hop = 0; hop < S->n - 1; ++hop) {
    int vals[3] = {good, 1, 1};
    MB_CUDA(c, cudaMemcpyAsync(flagdev, vals, sizeof(vals), cudaMemcpyHostToDevice, c->stream));
    MB_NCCL(c, api.GroupStart());
    if (lower >= 0) MB_NCCL(c, api.Send(flagdev, sizeof(int), ncclUint8, lower, S->comm, c->stream));
    if (upper >= 0) MB_NCCL(c, api.Send(flagdev, sizeof(int), ncclUint8, upper, S->comm, c->stream));
    if (upper >= 0) MB_NCCL(c, api.Recv(flagdev + 2, sizeof(int), ncclUint8, upper, S->comm, c->stream));
    if (lower >= 0) MB_NCCL(c, api.Recv(flagdev + 1, sizeof(int), ncclUint8, lower, S->comm, c->stream));
    MB_NCCL(c, api.GroupEnd());
    MB_CUDA(c, cudaMemcpyAsync(vals, flagdev, sizeof(vals), cudaMemcpyDeviceToHost, c->stream));
    MB_CUDA(c, cudaStreamSynchronize(c->stream));
    good = good && vals[1] && vals[2];
  }
  if (!good) return MB_OK;
  out->on = true;
  out->lower = lo;
  out->upper = up;
  return MB_OK;
}

static int stream_wait_geq(Ctx *c, const unsigned *flag, unsigned value) {
  ShardState *S = c->shard;
  if (S->wait_value) {
    if (S->wait_value((CUstream)c->stream, (CUdeviceptr)flag, value, CU_STREAM_WAIT_VALUE_GEQ) == CUDA_SUCCESS) return MB_OK;
    S->wait_value = nullptr;  // not supported here: fall through to the spinning kernel from now on
  }
  spin_flag_kernel<<<1, 1, 0, c->stream>>>(flag, value);
  return check_cuda(c, cudaGetLastError(), "spin_flag_kernel launch");
}

// copy my boundary planes of `buf` (parity bidx) into the neighbours' staging slots and tell them.
// The slot was last used for round - 2, so the neighbour must have copied that round out (credit_need).
static int send_ghosts_p2p(Ctx *c, const mb_shard_plan &sp, const SlabGeom &g, const P2P &pp, char *const bufs[2], int bidx,
                           unsigned round, unsigned credit_need) {
  ShardState *S = c->shard;
  char *buf = bufs[bidx];
  if (sp.send_first_to >= 0) {  // my first halo_hi planes -> the lower neighbour's UPPER ghost planes
    int st = stream_wait_geq(c, &S->flags[F_CREDIT_LO], credit_need);
    if (st) return st;
    MB_CUDA(c, cudaMemcpyAsync(pp.lower.stage + stage_slot(g, bidx, 1), buf + (size_t)g.lo * g.plane_bytes,
                               (size_t)g.hi * g.plane_bytes, cudaMemcpyDeviceToDevice, c->stream));
    set_flag_kernel<<<1, 1, 0, c->stream>>>(&pp.lower.flags[F_DELIV_HI], round);
  }
  if (sp.send_last_to >= 0) {  // my last halo_lo planes -> the upper neighbour's LOWER ghost planes
    int st = stream_wait_geq(c, &S->flags[F_CREDIT_HI], credit_need);
    if (st) return st;
    MB_CUDA(c, cudaMemcpyAsync(pp.upper.stage + stage_slot(g, bidx, 0), buf + (size_t)g.local * g.plane_bytes,
                               (size_t)g.lo * g.plane_bytes, cudaMemcpyDeviceToDevice, c->stream));
    set_flag_kernel<<<1, 1, 0, c->stream>>>(&pp.upper.flags[F_DELIV_LO], round);
  }
  return check_cuda(c, cudaGetLastError(), "set_flag_kernel launch");
}

// wait for this round's staged planes, move them into the ghost planes of `buf` (parity bidx) and hand
// the staging slots back to the neighbours
static int recv_ghosts_p2p(Ctx *c, const mb_shard_plan &sp, const SlabGeom &g, const P2P &pp, char *const bufs[2], int bidx,
                           unsigned round) {
  ShardState *S = c->shard;
  char *buf = bufs[bidx];
  int st = MB_OK;
  if (sp.recv_lower_from >= 0) {
    if ((st = stream_wait_geq(c, &S->flags[F_DELIV_LO], round))) return st;
    MB_CUDA(c, cudaMemcpyAsync(buf, S->xstage + stage_slot(g, bidx, 0), (size_t)g.lo * g.plane_bytes, cudaMemcpyDeviceToDevice,
                               c->stream));
    set_flag_kernel<<<1, 1, 0, c->stream>>>(&pp.lower.flags[F_CREDIT_HI], round);
  }
  if (sp.recv_upper_from >= 0) {
    if ((st = stream_wait_geq(c, &S->flags[F_DELIV_HI], round))) return st;
    MB_CUDA(c, cudaMemcpyAsync(buf + (size_t)(g.lo + g.local) * g.plane_bytes, S->xstage + stage_slot(g, bidx, 1),
                               (size_t)g.hi * g.plane_bytes, cudaMemcpyDeviceToDevice, c->stream));
    set_flag_kernel<<<1, 1, 0, c->stream>>>(&pp.upper.flags[F_CREDIT_LO], round);
  }
  return check_cuda(c, cudaGetLastError(), "set_flag_kernel launch");
}

}  // namespace mb

using namespace mb;

extern "C" {

int mb_shard_plan_make(const mb_stencil_desc *desc, const int64_t *global_dims, int n_ranks, int rank,
                       mb_shard_plan *out) {
  return shard_plan(desc, global_dims, n_ranks, rank, out, nullptr);
}

const char *mb_shard_transport(mb_ctx *h) { return (h && h->c.shard) ? h->c.shard->transport : "none"; }

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
  // scratch for the IPC-handle exchange of the peer-memory path: allocated here, where a failure is reported to
  // the caller on every rank alike, so that setup_p2p can always take part in the exchange
  MB_CUDA(c, cudaMalloc(&c->shard->stage, 3 * sizeof(IpcInfo)));
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
  int tb2_mode = 0;  // 1: fused pairs under the caller's finite promise; 2: fused pairs, checked (k_vol3d_tb2.cu)
  if (extra_lo > 0 && extra_lo == sp.step_halo_lo && extra_hi == sp.step_halo_hi && sp.halo_lo == 2 && sp.halo_hi == 2) {
    in_dims[0] = sp.halo_lo + sp.local_planes + sp.halo_hi;
    if ((st = get_dev_plan(c, &ld, in_dims, &dp2))) return st;
    tb2_mode = vol3d_tb2_slab_mode(c, dp2->p);
    if (tb2_mode == 2 && ensure_flag(c) != MB_OK) tb2_mode = 0;
  }

  SlabGeom g;
  g.plane_bytes = (size_t)dp->p.esize;
  for (int i = 1; i < desc->rank; ++i) g.plane_bytes *= (size_t)local_dims[i];
  g.local = sp.local_planes; g.lo = sp.halo_lo; g.hi = sp.halo_hi;
  char *bufs[2] = {(char *)dev_a, (char *)dev_b};

  // how the halo planes travel: peer-memory copies when every rank could map its neighbours' staging
  // areas, NCCL send/recv otherwise (identical decision on every rank)
  P2P pp;
  if ((st = setup_p2p(c, sp, g, &pp))) return st;

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
          c->aux_launches++;
        }
    }
  MB_CUDA(c, cudaGetLastError());

  unsigned k = S->rounds;  // rounds completed on this communicator so far
  S->transport = pp.on ? "p2p" : (S->n > 1 ? "nccl" : "none");

  // ghosts of the initial state
  if ((st = fill_local_ghosts(c, sp, g, bufs[0], *dp, false, c->stream))) return st;
  if (pp.on) {
    if ((st = send_ghosts_p2p(c, sp, g, pp, bufs, 0, k + 1, k))) return st;
  } else {
    MB_CUDA(c, cudaEventRecord(c->ev_a, c->stream));
    MB_CUDA(c, cudaStreamWaitEvent(c->comm_stream, c->ev_a, 0));
    if ((st = exchange(c, sp, g, bufs[0], c->comm_stream))) return st;
    MB_CUDA(c, cudaEventRecord(c->ev_b, c->comm_stream));
    MB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_b, 0));
  }

  const bool talks = sp.recv_lower_from >= 0 || sp.recv_upper_from >= 0 || sp.send_first_to >= 0 || sp.send_last_to >= 0;
  int cur = 0;
  for (int64_t s = 0; s < n_steps;) {
    const bool two = tb2_mode && n_steps - s >= 2;  // a fused pair of steps, or one step
    const bool checked = two && tb2_mode == 2;       // no finite promise: nothing may leave before the check
    const int64_t adv = two ? 2 : 1;
    char *in = bufs[cur], *out_base = bufs[1 - cur];
    char *out = out_base + (size_t)g.lo * g.plane_bytes;  // the rank's own planes start after the lower ghost
    const bool last = s + adv == n_steps;
    ++k;  // this round's number (the same on every rank)
    auto launch = [&](const OuterRange &r) -> int {
      if (r.o1 <= r.o0) return MB_OK;
      if (two && !checked) return launch_vol3d_tb2(c, dp2->p, in, out, &r, (int)g.lo, sp.first_plane - g.lo, global_dims[0]);
      if (two) {
        // optimistic pair over the whole slab; if the kernel met Inf / NaN the pair is redone with the strict
        // kernel: in -> (intermediate, parked in the out buffer one plane up) -> own planes of in -> out.
        // The four follow-up launches return at once while the flag is clear.
        MB_CUDA(c, cudaMemsetAsync(c->d_flag, 0, sizeof(unsigned), c->stream));
        int e = launch_vol3d_tb2(c, dp2->p, in, out, &r, (int)g.lo, sp.first_plane - g.lo, global_dims[0], c->d_flag);
        if (e) return e;
        char *mid = out_base + g.plane_bytes;  // intermediate planes -1 .. local of the slab
        if ((e = launch_vol3d_dense_cond(c, *dp2, in, mid, c->d_flag))) return e == -1 ? fail(c, MB_ERR_UNSUPPORTED, "internal: strict redo") : e;
        const size_t plane_elems = g.plane_bytes / dp->p.esize;
        if (sp.first_plane - 1 < 0)  // the border applies to the intermediate array: Fill outside the global array
          if ((e = launch_cond_fill(c, mid, plane_elems, dp->p.esize, 0ull, c->d_flag))) return e;
        if (sp.first_plane + g.local >= global_dims[0])
          if ((e = launch_cond_fill(c, mid + (size_t)(g.local + 1) * g.plane_bytes, plane_elems, dp->p.esize, 0ull, c->d_flag))) return e;
        char *own_in = in + (size_t)g.lo * g.plane_bytes;
        if ((e = launch_vol3d_dense_cond(c, *dp, mid, own_in, c->d_flag))) return e == -1 ? fail(c, MB_ERR_UNSUPPORTED, "internal: strict redo") : e;
        if ((e = launch_cond_copy(c, out, own_in, (size_t)g.local * g.plane_bytes, c->d_flag))) return e;
        c->last_kernel = "vol3d_tb2";
        return MB_OK;
      }
      return launch_plan(c, *dp, in + (size_t)extra_lo * g.plane_bytes, out, &r);
    };
    if (pp.on && (st = recv_ghosts_p2p(c, sp, g, pp, bufs, cur, k))) return st;  // the neighbours' planes for this round
    // 1. the planes the neighbours are waiting for (a full ghost depth at either end) first
    int64_t top_end = g.hi < g.local ? g.hi : g.local;                        // result planes [0, top_end)
    int64_t bot_begin = g.local - g.lo > top_end ? g.local - g.lo : top_end;  // result planes [bot_begin, local)
    if (last || !talks || checked) { top_end = 0; bot_begin = g.local; }  // nobody waits / may not be served yet: one launch
    const OuterRange r1{0, top_end}, r2{bot_begin, g.local}, r3{top_end, bot_begin};
    if (checked) {
      if ((st = launch(r3))) return st;
    } else {
      if ((st = launch(r1))) return st;
      if ((st = launch(r2))) return st;
    }
    if (!last) {
      if (pp.on) {
        // 2. copy engines push the new boundary planes into the neighbours' staging slots of that parity;
        //    the neighbour must have copied round k-1's planes out of them
        if ((st = send_ghosts_p2p(c, sp, g, pp, bufs, 1 - cur, k + 1, k - 1))) return st;
      } else {
        MB_CUDA(c, cudaEventRecord(c->ev_a, c->stream));
        MB_CUDA(c, cudaStreamWaitEvent(c->comm_stream, c->ev_a, 0));
        // 2. neighbour exchange of the new boundary planes on the second stream, concurrent with
        if ((st = exchange(c, sp, g, out_base, c->comm_stream))) return st;
        MB_CUDA(c, cudaEventRecord(c->ev_b, c->comm_stream));
      }
    }
    // 3. the interior
    if (!checked && (st = launch(r3))) return st;
    if (!last) {
      // physical-border ghosts that are copies of own planes (Edge / Reflect / Continue / 1-rank Wrap)
      if ((st = fill_local_ghosts(c, sp, g, out_base, *dp, false, c->stream))) return st;
      // 4. join
      if (!pp.on) MB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_b, 0));
    }
    cur = 1 - cur;
    s += adv;
  }
  if (pp.on) S->rounds = k;
  *result = bufs[cur];
  return MB_OK;
}
