// Internal declarations shared by the translation units of libmassiv_b200.so.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/massiv_b200.h"

namespace mb {

// Host-side normal form of one descriptor applied to one input shape.  Everything the kernels
// need is derived here once: geometry (Array/Stencil.hs:197-217), the ordered tap list
// (fold: Stencil.hs:298-302; FromKernel: Stencil/Convolution.hs:64-80,107-121) and the
// coefficient bytes.
struct Plan {
  int kind = 0, elem = 0, rank = 0, border = 0;
  uint32_t flags = 0;
  int esize = 0;
  int64_t in_dims[MB_MAX_RANK] = {0};
  int64_t full_dims[MB_MAX_RANK] = {0};  // output extent before the stride is applied
  int64_t out_dims[MB_MAX_RANK] = {0};   // strideSize of full_dims
  int64_t stride[MB_MAX_RANK] = {0};
  int64_t shift[MB_MAX_RANK] = {0};      // source index of output cell 0: centre - pad_lo
  int64_t size[MB_MAX_RANK] = {0};       // tap window (desc.size)
  int64_t geom[MB_MAX_RANK] = {0};       // size used for the geometry (AVG: max(size,1))
  int64_t center[MB_MAX_RANK] = {0};
  int64_t pad_lo[MB_MAX_RANK] = {0}, pad_hi[MB_MAX_RANK] = {0};
  int64_t min_off[MB_MAX_RANK] = {0}, max_off[MB_MAX_RANK] = {0};
  int64_t ntaps = 0;
  int64_t avg_n = 0;                     // totalElem of desc.size
  std::vector<int32_t> tap_off;          // ntaps * rank, application order
  std::vector<uint8_t> c1, c2;           // ntaps coefficients each (elem bytes), may be empty
  uint8_t fill[8] = {0};
  std::vector<mb_post_op> post;          // fmap'ed functions applied to every result, in order
  // MB_EXPR: the sub-stencils ("terms") whose values the postfix program combines.  Their taps and coefficients
  // are concatenated in tap_off / c1 (term t owns [begin, begin + ntaps)); ntaps is the total.
  struct Term {
    int32_t kind;      // mb_kind of the term
    int32_t begin;     // first tap in tap_off / c1
    int32_t ntaps;
    int32_t avg_n;     // AVG: totalElem of the term's size
  };
  std::vector<Term> terms;
  std::vector<mb_xins> prog;
  int64_t in_elems = 0, out_elems = 0;
  bool dense = false;                    // taps form the full [0,size) window (fold/conv/corr/mag2)
  bool descending = false;               // dense taps visited with descending offsets (conv)
  bool same_shape() const {
    for (int i = 0; i < rank; ++i) if (out_dims[i] != in_dims[i]) return false;
    return true;
  }
  bool unit_stride() const {
    for (int i = 0; i < rank; ++i) if (stride[i] != 1) return false;
    return true;
  }
};

// validation + geometry only (no tap list): mb_out_dims
int plan_geometry(const mb_stencil_desc *d, const int64_t *in_dims, Plan *p, std::string *err);
// full plan
int make_plan(const mb_stencil_desc *d, const int64_t *in_dims, Plan *p, std::string *err);
int border_index(int border, int64_t k, int64_t i, int64_t *out, int32_t *is_fill);

// A plan whose tap table / coefficients are resident on the device.
struct DevPlan {
  Plan p;
  int32_t *d_taps = nullptr;
  int64_t *d_tap_lin = nullptr;  // the taps as linear offsets in the input (cells whose window is inside the array)
  void *d_c1 = nullptr, *d_c2 = nullptr;
  mutable void *d_sw = nullptr;  // small-window kernel: doubled tap list + both kernels of a MAG2 plan (made on first use)
};

struct ShardState;  // mb_shard.cu

struct Ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr, comm_stream = nullptr, out_stream = nullptr;
  std::vector<cudaEvent_t> pipe_events;  // host->host pipeline (mb_run): created on first use
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_a = nullptr, ev_b = nullptr;
  std::mutex mu;
  std::string err;
  bool post_fused = false;   // set by a kernel family that applied the plan's post-maps in its store epilogue
  int64_t launches = 0;
  int64_t aux_launches = 0;  // of those: conditional follow-ups of the checked kernels, post-map passes, ghost fills
  const char *last_kernel = "none";
  int num_sms = 148;
  int force_generic = 0;   // mb_ctx_set_option("force_generic")
  int life_bitpack = 1;    // mb_ctx_set_option("life_bitpack")
  int no_tma = 0;          // mb_ctx_set_option("no_tma"): tiles are filled with ordinary loads
  int no_known = 0;        // mb_ctx_set_option("no_known"): skip the compile-time-coefficient kernels
  int tb2 = 1;             // mb_ctx_set_option("tb2"): fuse pairs of steps of iterated 3-D star stencils
  int64_t tile_min_elems = 16384;  // below this many output cells the generic kernel is used
  int no_optimistic = 0;           // 1: never skip zero taps without the caller's promise (always the 27-tap kernel)
  int pipeline = 1;                // mb_run overlaps H2D / kernel / D2H over chunks of the outer dimension
  int prefer_smallwin = 0;         // 1: rank-2 plans try the small-window kernel before the compile-time-shaped ones (measurements)
  int no_p2p = 0;                  // sharded iteration: halo planes by ncclSend/ncclRecv instead of peer-memory copies
  std::map<std::string, std::shared_ptr<DevPlan>> plans;  // keyed by descriptor + dims bytes
  // grow-only device scratch used by the host->host entry points and two-pass fallbacks
  unsigned *d_flag = nullptr;  // "the optimistic kernel met a non-finite value": set by it, read by the conditional redo
  unsigned long long *d_conv = nullptr;  // convergence verdict of iterate_until (k_converge.cu)
  void *scratch[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // (3, 4: the pitched pair of mb_iterate_dev)
  size_t scratch_bytes[5] = {0, 0, 0, 0, 0};
  // Iterating on internally PADDED arrays (mb_api.cu, iterate_pitch): the row pitch in elements of the input / output
  // array the rank-3 kernels see while set; 0 = that array is dense.  Only vol3d / vol3d_tb2 launches happen meanwhile.
  int64_t in_pitch = 0, out_pitch = 0;
  size_t wavefront_min_bytes = (size_t)256 << 20;  // mb_iterate hides its copies behind the passes from this size on ("wavefront_min_mb")
  int no_pitch = 0;        // mb_ctx_set_option("no_pitch"): iterations stay in the caller's layout
  void *pinned = nullptr;
  size_t pinned_bytes = 0;
  ShardState *shard = nullptr;
  // Life (bit-packed, temporally blocked): boundary-row mailboxes and per-CTA progress flags
  unsigned long long *life_bound = nullptr;
  size_t life_bound_words = 0;
  int *life_flags = nullptr;
  int life_nflags = 0;
};

int fail(Ctx *c, int status, const std::string &msg);
int check_cuda(Ctx *c, cudaError_t e, const char *what);
#define MB_CUDA(ctx, expr)                                          \
  do {                                                              \
    int _st = ::mb::check_cuda((ctx), (expr), #expr);               \
    if (_st) return _st;                                            \
  } while (0)

int get_dev_plan(Ctx *c, const mb_stencil_desc *d, const int64_t *in_dims, std::shared_ptr<DevPlan> *out);
int ensure_scratch(Ctx *c, int slot, size_t bytes, void **ptr);

// An output window along the outermost dimension: only output planes [o0, o1) are produced
// (used by the slab-sharded path to peel boundary planes from the interior).
struct OuterRange {
  int64_t o0, o1;
};

// dispatcher: picks the kernel family for the plan and launches it on c->stream
int launch_plan(Ctx *c, const DevPlan &dp, const void *d_in, void *d_out, const OuterRange *range);

// kernel families (each returns MB_OK, an error, or -1 for "not applicable, try the next one")
int launch_linescan(Ctx *c, const DevPlan &dp, const void *d_in, void *d_out, const OuterRange *range);  // -1: not applicable
int launch_direct2d(Ctx *c, const DevPlan &dp, const void *d_in, void *d_out, const OuterRange *range);  // -1: not applicable
int launch_post(Ctx *c, const Plan &p, void *d_out, const OuterRange *range);  // in-place post-maps over the output
int ensure_flag(Ctx *c);  // allocates c->d_flag
// iterateUntil predicates: cells with !(cur == prev) and the largest |cur - prev| (blocks until known)
int converge_check(Ctx *c, int elem, const void *cur, const void *prev, size_t n, unsigned long long *unequal, double *max_abs_diff);
// strict 27-tap step that runs only when *cond != 0 (nullptr: always)
int launch_vol3d_dense_cond(Ctx *c, const DevPlan &dp, const void *d_in, void *d_out, const unsigned *cond);
// dst <- src when *cond != 0
int launch_cond_copy(Ctx *c, void *dst, const void *src, size_t bytes, const unsigned *cond);
// 0: two-step passes not possible; 1: under the caller's finite promise; 2: optimistic (checked) — see k_vol3d_tb2.cu
int vol3d_tb2_mode(Ctx *c, const Plan &p);
int launch_generic(Ctx *c, const DevPlan &dp, const void *d_in, void *d_out, const OuterRange *range);
int launch_tile2d(Ctx *c, const DevPlan &dp, const void *d_in, void *d_out, const OuterRange *range);
// rank 2, unit stride, any window up to 16 x 16, every kind incl. MB_TAPS / MB_EXPR, every element type (k_smallwin.cu)
int launch_smallwin(Ctx *c, const DevPlan &dp, const void *d_in, void *d_out, const OuterRange *range);
int launch_vol3d(Ctx *c, const DevPlan &dp, const void *d_in, void *d_out, const OuterRange *range);
int launch_life(Ctx *c, const DevPlan &dp, const void *d_in, void *d_out);
// Life: many generations in one go (temporal blocking); returns -1 when not applicable
int life_iterate(Ctx *c, const DevPlan &dp, void *d_a, void *d_b, int64_t n_steps, void **result);
// 3-D star stencils, two steps per pass (k_vol3d_tb2.cu)
bool vol3d_tb2_applicable(Ctx *c, const Plan &p);
int vol3d_tb2_slab_mode(Ctx *c, const Plan &p);
int launch_cond_fill(Ctx *c, void *dst, size_t n_elems, int esize, unsigned long long bits, const unsigned *cond);
bool tb2_desc_eligible(const mb_stencil_desc *d);  // host-only: could this descriptor ever take the fused path?
int launch_vol3d_tb2(Ctx *c, const Plan &p, const void *d_in, void *d_out, const OuterRange *range, int sz2, int64_t zoff,
                     int64_t dglob, unsigned *detect = nullptr);

void shard_destroy(Ctx *c);
int shard_plan(const mb_stencil_desc *d, const int64_t *global_dims, int n_ranks, int rank, mb_shard_plan *out,
               std::string *err);

}  // namespace mb

struct mb_ctx {
  mb::Ctx c;
};
