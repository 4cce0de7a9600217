// C-ABI entry points of libmassiv_b200.so (declared in include/massiv_b200.h): context, memory,
// plan cache, kernel dispatch, run / iterate.  No torch types, no C++ exceptions across the boundary.
#include <sched.h>

#include <cstdio>
#include <cstdlib>
#include <functional>
#include <new>

#include "mb_internal.h"

namespace mb {

int fail(Ctx *c, int status, const std::string &msg) {
  if (c) c->err = msg;
  return status;
}

int check_cuda(Ctx *c, cudaError_t e, const char *what) {
  if (e == cudaSuccess) return MB_OK;
  std::string m = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
  if (c) c->err = m;
  return e == cudaErrorMemoryAllocation ? MB_ERR_OOM : MB_ERR_CUDA;
}

int ensure_scratch(Ctx *c, int slot, size_t bytes, void **ptr) {
  if (bytes > c->scratch_bytes[slot]) {
    if (c->scratch[slot]) {
      MB_CUDA(c, cudaStreamSynchronize(c->stream));
      MB_CUDA(c, cudaFree(c->scratch[slot]));
      c->scratch[slot] = nullptr; c->scratch_bytes[slot] = 0;
    }
    MB_CUDA(c, cudaMalloc(&c->scratch[slot], bytes));
    c->scratch_bytes[slot] = bytes;
  }
  *ptr = c->scratch[slot];
  return MB_OK;
}

static void free_dev_plan(DevPlan *dp) {
  if (dp->d_taps) cudaFree(dp->d_taps);
  if (dp->d_tap_lin) cudaFree(dp->d_tap_lin);
  if (dp->d_c1) cudaFree(dp->d_c1);
  if (dp->d_c2) cudaFree(dp->d_c2);
  if (dp->d_sw) cudaFree(dp->d_sw);
  delete dp;
}

int get_dev_plan(Ctx *c, const mb_stencil_desc *d, const int64_t *in_dims, std::shared_ptr<DevPlan> *out) {
  Plan p;
  std::string err;
  int st = make_plan(d, in_dims, &p, &err);
  if (st) return fail(c, st, err);
  // cache key: every byte that determines the plan
  std::string key;
  key.append((const char *)d, offsetof(mb_stencil_desc, coeffs));
  key.append((const char *)&d->ntaps, sizeof(d->ntaps));
  key.append((const char *)&d->flags, sizeof(d->flags));
  key.append((const char *)in_dims, sizeof(int64_t) * d->rank);
  key.append((const char *)p.tap_off.data(), p.tap_off.size() * sizeof(int32_t));
  key.append((const char *)p.c1.data(), p.c1.size());
  key.append((const char *)p.c2.data(), p.c2.size());
  key.append((const char *)p.post.data(), p.post.size() * sizeof(mb_post_op));
  key.append((const char *)p.terms.data(), p.terms.size() * sizeof(Plan::Term));
  key.append((const char *)p.prog.data(), p.prog.size() * sizeof(mb_xins));
  auto it = c->plans.find(key);
  if (it != c->plans.end()) { *out = it->second; return MB_OK; }
  if (c->plans.size() > 256) c->plans.clear();  // bounded; shared_ptrs keep in-flight plans alive
  std::shared_ptr<DevPlan> dp(new DevPlan(), free_dev_plan);
  dp->p = std::move(p);
  const Plan &q = dp->p;
  if (!q.tap_off.empty()) {
    MB_CUDA(c, cudaMalloc(&dp->d_taps, q.tap_off.size() * sizeof(int32_t)));
    MB_CUDA(c, cudaMemcpy(dp->d_taps, q.tap_off.data(), q.tap_off.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  if (!q.tap_off.empty()) {
    std::vector<int64_t> lin((size_t)q.ntaps);
    for (int64_t t = 0; t < q.ntaps; ++t) {
      int64_t l = 0;
      for (int i = 0; i < q.rank; ++i) l = l * q.in_dims[i] + q.tap_off[(size_t)t * q.rank + i];
      lin[(size_t)t] = l;
    }
    MB_CUDA(c, cudaMalloc(&dp->d_tap_lin, lin.size() * sizeof(int64_t)));
    MB_CUDA(c, cudaMemcpy(dp->d_tap_lin, lin.data(), lin.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
  }
  if (!q.c1.empty()) {
    MB_CUDA(c, cudaMalloc(&dp->d_c1, q.c1.size()));
    MB_CUDA(c, cudaMemcpy(dp->d_c1, q.c1.data(), q.c1.size(), cudaMemcpyHostToDevice));
  }
  if (!q.c2.empty()) {
    MB_CUDA(c, cudaMalloc(&dp->d_c2, q.c2.size()));
    MB_CUDA(c, cudaMemcpy(dp->d_c2, q.c2.data(), q.c2.size(), cudaMemcpyHostToDevice));
  }
  // the tables are uploaded synchronously: a cached plan may later be launched on another stream
  // (mb_ctx_set_stream) with no ordering against the stream that was current here
  c->plans[key] = dp;
  *out = dp;
  return MB_OK;
}

// Kernel-family dispatch.  Order: most specialised first; each family answers -1 when the plan is
// outside what it covers.  The generic kernel covers everything.
static int launch_family(Ctx *c, const DevPlan &dp, const void *d_in, void *d_out, const OuterRange *range);

int launch_plan(Ctx *c, const DevPlan &dp, const void *d_in, void *d_out, const OuterRange *range) {
  if (dp.p.out_elems == 0) return MB_OK;
  c->post_fused = false;
  int st = launch_family(c, dp, d_in, d_out, range);
  if (st || dp.p.post.empty() || c->post_fused) return st;
  // post-maps run as an in-place pass over what the stencil kernel just wrote (same stream)
  const char *family = c->last_kernel;
  st = launch_post(c, dp.p, d_out, range);
  c->last_kernel = family;
  return st;
}

// Arrays of rank 3-5 whose stencil has extent 1 in the leading dimensions (a 2-D stencil applied to every
// slice of a stack, a 3-D one to every volume of a batch): the tiled rank-2 / rank-3 kernels run once per
// slice on a reduced plan.  -1 (and nothing launched) when the shape is not of that kind or the tiled
// kernels would not take the slices.
static int launch_batched(Ctx *c, const DevPlan &dp, const void *d_in, void *d_out, const OuterRange *range) {
  const Plan &p = dp.p;
  if (p.rank < 3 || p.kind == MB_TAPS || p.kind == MB_LIFE || p.kind == MB_EXPR) return -1;
  int lead = 0;
  while (lead < p.rank && p.size[lead] == 1 && p.geom[lead] == 1 && p.stride[lead] == 1 && p.shift[lead] == 0 &&
         p.out_dims[lead] == p.in_dims[lead] && p.pad_lo[lead] == 0 && p.pad_hi[lead] == 0) ++lead;
  if (lead > p.rank - 2) lead = p.rank - 2;
  const int inner = p.rank - lead;
  if (lead < 1 || inner > 3) return -1;
  mb_stencil_desc rd;
  std::memset(&rd, 0, sizeof(rd));
  rd.kind = p.kind; rd.elem = p.elem; rd.rank = inner; rd.border = p.border; rd.flags = p.flags;
  int64_t in_slice = 1, out_slice = 1, batch = 1, per_outer = 1;
  for (int i = 0; i < lead; ++i) { batch *= p.in_dims[i]; if (i > 0) per_outer *= p.in_dims[i]; }
  for (int i = 0; i < inner; ++i) {
    const int j = lead + i;
    rd.size[i] = p.size[j]; rd.center[i] = p.center[j]; rd.pad_lo[i] = p.pad_lo[j]; rd.pad_hi[i] = p.pad_hi[j];
    rd.stride[i] = p.stride[j];
    in_slice *= p.in_dims[j]; out_slice *= p.out_dims[j];
  }
  std::memcpy(rd.fill, p.fill, 8);
  rd.coeffs = p.c1.empty() ? nullptr : p.c1.data();
  rd.coeffs2 = p.c2.empty() ? nullptr : p.c2.data();
  if (out_slice < c->tile_min_elems || batch > 65536 || batch < 1) return -1;
  std::shared_ptr<DevPlan> sub;
  if (get_dev_plan(c, &rd, &p.in_dims[lead], &sub) != MB_OK) return -1;
  int64_t b0 = 0, b1 = batch;
  if (range) { b0 = range->o0 * per_outer; b1 = range->o1 * per_outer; }
  const size_t ib = (size_t)in_slice * p.esize, ob = (size_t)out_slice * p.esize;
  for (int64_t b = b0; b < b1; ++b) {
    const char *in = (const char *)d_in + (size_t)b * ib;
    char *out = (char *)d_out + (size_t)b * ob;
    int st = inner == 2 ? launch_tile2d(c, *sub, in, out, nullptr) : launch_vol3d(c, *sub, in, out, nullptr);
    if (st == -1 && inner == 2 && p.post.empty()) st = launch_smallwin(c, *sub, in, out, nullptr);
    if (st == -1) {
      if (b == b0) return -1;  // not a shape the tiled kernels cover: nothing has been launched
      st = launch_generic(c, *sub, in, out, nullptr);
    }
    if (st) return st;
  }
  return MB_OK;
}

static int launch_family(Ctx *c, const DevPlan &dp, const void *d_in, void *d_out, const OuterRange *range) {
  if (!c->force_generic) {
    int st = range ? -1 : launch_life(c, dp, d_in, d_out);  // the Life kernels produce whole arrays
    if (st != -1) return st;
    // Mapped stencils (fmap f) of small windows: the small-window kernel applies the maps in its store epilogue;
    // behind the compile-time-shaped kernels they are a second pass over the output.  Measured (scripts/probe_post.py,
    // 16384^2): conv 3x3 |.|*c Float 384 vs 218 Gcells/s, avg 3x3 + c Double 218 vs 123, conv 5x5 sqrt|.| Float 206 vs 181;
    // without maps the shaped kernels win (563 vs 538, 605 vs 504, 396 vs 259).
    // (rank 2 only: the rank-3 small-window path re-stages a box per output plane and loses to vol3d + a post pass)
    if (c->prefer_smallwin || (!dp.p.post.empty() && dp.p.ntaps <= 32 && dp.p.rank == 2)) {
      st = launch_smallwin(c, dp, d_in, d_out, range);
      if (st != -1) return st;
    }
    st = launch_tile2d(c, dp, d_in, d_out, range);
    if (st != -1) return st;
    st = launch_vol3d(c, dp, d_in, d_out, range);
    if (st != -1) return st;
    st = launch_linescan(c, dp, d_in, d_out, range);
    if (st != -1) return st;
    st = launch_direct2d(c, dp, d_in, d_out, range);
    if (st != -1) return st;
    st = launch_batched(c, dp, d_in, d_out, range);  // (stacks of slices keep the rank-2 kernels' speed)
    if (st != -1) return st;
    st = launch_smallwin(c, dp, d_in, d_out, range);
    if (st != -1) return st;
  }
  return launch_generic(c, dp, d_in, d_out, range);
}

struct Guard {
  Ctx *c;
  std::unique_lock<std::mutex> lk;
  int st;
  explicit Guard(mb_ctx *h) : c(h ? &h->c : nullptr), st(MB_OK) {
    if (!c) { st = MB_ERR_INVALID_DESC; return; }
    lk = std::unique_lock<std::mutex>(c->mu);
    st = check_cuda(c, cudaSetDevice(c->device), "cudaSetDevice");  // no thread-local CUDA state assumed
  }
};

static int copy_h2d(Ctx *c, void *dst, const void *src, size_t bytes) {
  if (!bytes) return MB_OK;
  MB_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  return MB_OK;
}
static int copy_d2h(Ctx *c, void *dst, const void *src, size_t bytes) {
  if (!bytes) return MB_OK;
  MB_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  return MB_OK;
}

// One pair of steps with the two-step kernel cur -> nxt.  No promise given (flag): if the kernel met Inf / NaN (in
// the input or the intermediate array) the pair is redone with the strict kernel.  All redo launches return at once
// while the flag is clear.  With a spare array of nxt's... (see callers) the redo goes cur -> spare -> nxt; without,
// cur -> nxt -> cur and a copy to where the result belongs (cur and nxt then share a layout of `bytes` bytes).
static int tb2_pair(Ctx *c, const DevPlan *dp, int tb2, void *cur, void *nxt, void *spare, int64_t spare_pitch, size_t bytes) {
  unsigned *flag = tb2 == 2 ? c->d_flag : nullptr;
  if (flag) MB_CUDA(c, cudaMemsetAsync(flag, 0, sizeof(unsigned), c->stream));
  int st = launch_vol3d_tb2(c, dp->p, cur, nxt, nullptr, 0, 0, dp->p.in_dims[0], flag);
  if (st) return st;  // (-1: the array cannot be described to TMA and has no other staging path: single steps)
  if (flag) {
    if (spare) {
      const int64_t ip = c->in_pitch, op = c->out_pitch;
      c->out_pitch = spare_pitch;
      st = launch_vol3d_dense_cond(c, *dp, cur, spare, flag);
      c->in_pitch = spare_pitch; c->out_pitch = op;
      if (st == 0) st = launch_vol3d_dense_cond(c, *dp, spare, nxt, flag);
      c->in_pitch = ip;
      if (st == -1) return fail(c, MB_ERR_UNSUPPORTED, "internal: strict redo");
      if (st) return st;
    } else {
      if ((st = launch_vol3d_dense_cond(c, *dp, cur, nxt, flag)) == -1) return fail(c, MB_ERR_UNSUPPORTED, "internal: strict redo");
      if (st) return st;
      if ((st = launch_vol3d_dense_cond(c, *dp, nxt, cur, flag))) return st;
      if ((st = launch_cond_copy(c, nxt, cur, bytes, flag))) return st;
    }
    c->last_kernel = "vol3d_tb2";
  }
  return MB_OK;
}

// the iteration proper on two arrays of the same layout: dense, or (c->in_pitch = c->out_pitch set by the caller) padded
static int iterate_core(Ctx *c, const DevPlan *dp, void *a, void *b, int64_t n_steps, void **result) {
  int st;
  void *cur = a, *nxt = b;
  int64_t s = 0;
  const int tb2 = dp->p.same_shape() ? vol3d_tb2_mode(c, dp->p) : 0;
  if (tb2 == 2 && ensure_flag(c) != MB_OK) return fail(c, MB_ERR_CUDA, "cannot allocate the check flag");
  const size_t bytes = c->out_pitch ? (size_t)(dp->p.out_elems / dp->p.out_dims[dp->p.rank - 1]) * (size_t)c->out_pitch * dp->p.esize
                                    : (size_t)dp->p.out_elems * dp->p.esize;
  if (tb2) {
    // temporal blocking: pairs of steps in one pass over HBM; an odd last step runs singly below
    for (; s + 2 <= n_steps; s += 2) {
      st = tb2_pair(c, dp, tb2, cur, nxt, nullptr, 0, bytes);
      if (st == -1) break;
      if (st) return st;
      void *t = cur; cur = nxt; nxt = t;
    }
  }
  for (; s < n_steps; ++s) {
    st = launch_plan(c, *dp, cur, nxt, nullptr);
    if (st) return st;
    void *t = cur; cur = nxt; nxt = t;
  }
  *result = cur;
  return MB_OK;
}

// The same from DENSE a to DENSE b through a padded pair p1, p2 (row pitch `pitch`): the first pass reads the dense
// array and writes padded rows, the last one reads padded rows and writes b, the passes in between run on the pair.
static int iterate_through_padded(Ctx *c, const DevPlan *dp, void *a, void *b, void *p1, void *p2, int64_t pitch, int64_t n_steps) {
  const int tb2 = vol3d_tb2_mode(c, dp->p);
  if (tb2 == 2 && ensure_flag(c) != MB_OK) return fail(c, MB_ERR_CUDA, "cannot allocate the check flag");
  const size_t pbytes = (size_t)(dp->p.out_elems / dp->p.out_dims[2]) * (size_t)pitch * dp->p.esize;
  const int64_t pairs = n_steps / 2, npass = pairs + (n_steps & 1);
  void *cur = a;
  int64_t cur_pitch = 0;
  int st = MB_OK;
  for (int64_t i = 0; i < npass && st == MB_OK; ++i) {
    const bool last = i == npass - 1;
    void *out = last ? b : (cur == p1 ? p2 : p1);
    const int64_t out_pitch = last ? 0 : pitch;
    // a padded array that is neither read nor written by this pass (none while both of the pair are in use)
    void *spare = (cur != p1 && out != p1) ? p1 : ((cur != p2 && out != p2) ? p2 : nullptr);
    c->in_pitch = cur_pitch; c->out_pitch = out_pitch;
    if (i < pairs) {
      st = tb2_pair(c, dp, tb2, cur, out, spare, pitch, pbytes);
      if (st == -1) st = fail(c, MB_ERR_UNSUPPORTED, "internal: two-step kernel declined a padded pass");
    } else {
      st = launch_plan(c, *dp, cur, out, nullptr);
    }
    cur = out; cur_pitch = out_pitch;
  }
  c->in_pitch = c->out_pitch = 0;
  return st;
}

// Rank-3 arrays whose rows are not whole 16-byte units cannot be described to TMA, and the kernels' other staging paths
// run at two thirds of its speed.  An ITERATION does not have to live in the caller's layout: the two-step
// kernel's plans (3x3x3 star, Fill 0, same shape) iterate on a pair of arrays whose rows are padded to the next
// 16-byte multiple — the tensor map describes the logical width, so the pad column is never read or written.
// Returns the pitch in elements, 0 when the plan iterates where it is.
static int64_t iterate_pitch(Ctx *c, const DevPlan *dp, int64_t n_steps) {
  const Plan &p = dp->p;
  if (c->no_tma || c->force_generic || c->no_pitch || c->prefer_smallwin || p.rank != 3 || !p.same_shape() || !p.post.empty()) return 0;
  const int64_t w = p.in_dims[2], v = 16 / p.esize;
  if (p.esize < 4 || w % v == 0 || n_steps < 6 || p.in_elems < c->tile_min_elems) return 0;
  if (!vol3d_tb2_mode(c, p)) return 0;
  return (w + v - 1) / v * v;
}

static int copy_rows(Ctx *c, void *dst, size_t dpitch, const void *src, size_t spitch, size_t row_bytes, size_t rows, cudaMemcpyKind kind) {
  MB_CUDA(c, cudaMemcpy2DAsync(dst, dpitch, src, spitch, row_bytes, rows, kind, c->stream));
  return MB_OK;
}

static int iterate_dev_locked(Ctx *c, const mb_stencil_desc *desc, const int64_t *dims, void *a, void *b,
                              int64_t n_steps, void **result) {
  if (n_steps < 1) return fail(c, MB_ERR_INVALID_DESC, "iterateUntil applies the step at least once");
  std::shared_ptr<DevPlan> dp;
  int st = get_dev_plan(c, desc, dims, &dp);
  if (st) return st;
  if (!dp->p.same_shape()) return fail(c, MB_ERR_SIZE_MISMATCH, "iterate needs output size == input size");
  if (!c->force_generic && dp->p.post.empty()) {
    st = life_iterate(c, *dp, a, b, n_steps, result);
    if (st != -1) return st;
  }
  // the caller's arrays are dense: the first and the last pass change the layout, so the padded pair pays from three
  // passes on
  const int64_t pitch = iterate_pitch(c, dp.get(), n_steps);
  if (pitch) {
    const Plan &p = dp->p;
    const size_t pb = (size_t)(p.in_dims[0] * p.in_dims[1]) * (size_t)pitch * p.esize;
    void *p1 = nullptr, *p2 = nullptr;
    // (no room for the pair: iterate where the arrays are)
    if (ensure_scratch(c, 3, pb, &p1) == MB_OK && ensure_scratch(c, 4, pb, &p2) == MB_OK) {
      if ((st = iterate_through_padded(c, dp.get(), a, b, p1, p2, pitch, n_steps))) return st;
      *result = b;
      return MB_OK;
    }
    cudaGetLastError();
    c->err.clear();
  }
  return iterate_core(c, dp.get(), a, b, n_steps, result);
}

// iterateUntil with a device-side predicate: chunks of check_every applications; the last application of a chunk is
// a single step so that BOTH arrays the predicate compares exist (the fused two-step kernel never materialises the
// state in between).
static int iterate_until_locked(Ctx *c, const mb_stencil_desc *desc, const int64_t *dims, void *a, void *b, int64_t max_steps,
                                int64_t check_every, int predicate, double tol, int64_t *steps, int32_t *converged, void **result) {
  if (max_steps < 1 || check_every < 1) return fail(c, MB_ERR_INVALID_DESC, "iterateUntil applies the step at least once");
  if (predicate != MB_UNTIL_EQUAL && predicate != MB_UNTIL_MAX_ABS_DIFF) return fail(c, MB_ERR_INVALID_DESC, "unknown predicate");
  if (predicate == MB_UNTIL_MAX_ABS_DIFF && desc->elem != MB_F32 && desc->elem != MB_F64)
    return fail(c, MB_ERR_UNSUPPORTED, "the max-abs-difference predicate needs Float / Double elements");
  std::shared_ptr<DevPlan> dp;
  int st = get_dev_plan(c, desc, dims, &dp);
  if (st) return st;
  if (!dp->p.same_shape()) return fail(c, MB_ERR_SIZE_MISMATCH, "iterate needs output size == input size");
  void *cur = a, *other = b;
  int64_t done = 0;
  *converged = 0;
  while (done < max_steps) {
    const int64_t chunk = std::min<int64_t>(check_every, max_steps - done);
    if (chunk > 1) {
      void *r = nullptr;
      if ((st = iterate_dev_locked(c, desc, dims, cur, other, chunk - 1, &r))) return st;
      if (r != cur) { other = cur; cur = r; }
    }
    if ((st = launch_plan(c, *dp, cur, other, nullptr))) return st;  // prev = cur, cur = other
    void *prev = cur;
    cur = other; other = prev;
    done += chunk;
    unsigned long long unequal = 0;
    double maxd = 0.0;
    if ((st = converge_check(c, desc->elem, cur, prev, (size_t)dp->p.out_elems, &unequal, &maxd))) return st;
    const bool ok = predicate == MB_UNTIL_EQUAL ? unequal == 0 : maxd < tol;
    if (ok) { *converged = 1; break; }
  }
  *steps = done;
  *result = cur;
  return MB_OK;
}

static int check_sep2(Ctx *c, const mb_stencil_desc *f, const mb_stencil_desc *s) {
  if (!f || !s) return fail(c, MB_ERR_INVALID_DESC, "null descriptor");
  if (f->elem != s->elem || f->rank != s->rank || f->border != s->border ||
      std::memcmp(f->fill, s->fill, 8) != 0)
    return fail(c, MB_ERR_INVALID_DESC, "separable passes must share element type, rank and border");
  return MB_OK;
}

int run_sep2_fused(Ctx *c, const DevPlan &first, const DevPlan &second, const void *d_in, void *d_out, const OuterRange *range);

static int sep2_dev_locked(Ctx *c, const mb_stencil_desc *f, const mb_stencil_desc *s, const int64_t *dims,
                           const void *d_in, void *d_out, void *d_tmp) {
  int st = check_sep2(c, f, s);
  if (st) return st;
  std::shared_ptr<DevPlan> p1, p2;
  if ((st = get_dev_plan(c, f, dims, &p1))) return st;
  if ((st = get_dev_plan(c, s, dims, &p2))) return st;
  if (!p1->p.same_shape() || !p2->p.same_shape())
    return fail(c, MB_ERR_SIZE_MISMATCH, "separable passes must be mapStencil-shaped (same size in and out)");
  if (!c->force_generic && p1->p.post.empty() && p2->p.post.empty()) {
    st = run_sep2_fused(c, *p1, *p2, d_in, d_out, nullptr);
    if (st != -1) return st;
  }
  if (!d_tmp) {
    if ((st = ensure_scratch(c, 2, (size_t)p1->p.out_elems * p1->p.esize, &d_tmp))) return st;
  }
  if ((st = launch_plan(c, *p1, d_in, d_tmp, nullptr))) return st;
  return launch_plan(c, *p2, d_tmp, d_out, nullptr);
}

// Host -> host single pass as a three-stage pipeline over chunks of the outermost dimension: the input
// travels up in chunks on the copy stream, a chunk of output planes is computed as soon as the input
// planes it reads (its own plus the stencil's reach) have landed, and leaves on a third stream while the
// next chunk is computed.  PCIe is full duplex, so the two copies overlap each other as well as the
// kernel.  -1: not worth it / not applicable (small arrays, strided evaluation): the caller runs serially.
// `launch` (optional): what produces output planes [o0, o1) from the device input; default: the plan itself
int run_pipelined(Ctx *c, const DevPlan &dp, const void *host_in, void *host_out, void *d_in, void *d_out,
                  const std::function<int(const OuterRange &)> *launch = nullptr) {
  const Plan &p = dp.p;
  if (!c->pipeline || !p.unit_stride() || p.out_elems == 0 || p.in_elems == 0) return -1;
  const size_t ib = (size_t)p.in_elems * p.esize, ob = (size_t)p.out_elems * p.esize;
  const int64_t D = p.in_dims[0], OD = p.out_dims[0];
  int nch = (int)std::min<size_t>(16, (ib + ob) / ((size_t)96 << 20));
  const int64_t reach = std::max<int64_t>(1, std::max(-(p.shift[0] + p.min_off[0]), p.shift[0] + p.max_off[0]));
  while (nch > 1 && (OD / nch < 4 * (reach + 1) || D / nch < 4 * (reach + 1))) --nch;
  if (nch < 2) return -1;
  if (!c->out_stream && cudaStreamCreateWithFlags(&c->out_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return -1; }
  while ((int)c->pipe_events.size() < 2 * 16) {
    cudaEvent_t e;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return -1; }
    c->pipe_events.push_back(e);
  }
  size_t in_plane = (size_t)p.esize, out_plane = (size_t)p.esize;
  for (int i = 1; i < p.rank; ++i) { in_plane *= (size_t)p.in_dims[i]; out_plane *= (size_t)p.out_dims[i]; }
  // whatever the caller queued on the compute stream before comes first
  MB_CUDA(c, cudaEventRecord(c->ev_a, c->stream));
  MB_CUDA(c, cudaStreamWaitEvent(c->comm_stream, c->ev_a, 0));
  MB_CUDA(c, cudaStreamWaitEvent(c->out_stream, c->ev_a, 0));
  auto in_begin = [&](int j) { return D * j / nch; };
  // once anything has been queued, an error may not return while copies into the scratch buffers or into the
  // caller's host_out are still in flight (the caller may free or reuse them): drain the three streams first
  auto bail = [&](int st) {
    cudaStreamSynchronize(c->comm_stream);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->out_stream);
    cudaGetLastError();
    return st;
  };
#define MB_PIPE(expr)                                      \
  do {                                                     \
    int _ps = ::mb::check_cuda(c, (expr), #expr);          \
    if (_ps) return bail(_ps);                             \
  } while (0)
  for (int j = 0; j < nch; ++j) {
    const int64_t a = in_begin(j), b = in_begin(j + 1);
    MB_PIPE(cudaMemcpyAsync((char *)d_in + a * in_plane, (const char *)host_in + a * in_plane, (size_t)(b - a) * in_plane,
                               cudaMemcpyHostToDevice, c->comm_stream));
    MB_PIPE(cudaEventRecord(c->pipe_events[j], c->comm_stream));
  }
  for (int k = 0; k < nch; ++k) {
    const OuterRange r{OD * k / nch, OD * (k + 1) / nch};
    // the last input plane this chunk reads (index repaired by Edge / Reflect / Continue stays within
    // the stencil's reach of the array's ends; Wrap reaches the far end)
    int64_t need_hi = r.o1 - 1 + p.shift[0] + p.max_off[0];
    const int64_t need_lo = r.o0 + p.shift[0] + p.min_off[0];
    if (need_hi > D - 1) need_hi = D - 1;
    if (need_hi < 0) need_hi = 0;
    int j_hi = 0;
    while (j_hi + 1 < nch && in_begin(j_hi + 1) <= need_hi) ++j_hi;
    const bool leaves = need_lo < 0 || r.o1 - 1 + p.shift[0] + p.max_off[0] > D - 1;
    if (leaves && p.border != MB_FILL && p.border != MB_EDGE) {
      // repaired indices: Wrap lands at the opposite end; Reflect / Continue within `reach` of this end
      j_hi = p.border == MB_WRAP ? nch - 1 : std::max(j_hi, 0);
      if (p.border != MB_WRAP) {
        const int64_t far = std::min<int64_t>(D - 1, std::max<int64_t>(need_hi, reach + 1));
        while (j_hi + 1 < nch && in_begin(j_hi + 1) <= far) ++j_hi;
      }
    }
    MB_PIPE(cudaStreamWaitEvent(c->stream, c->pipe_events[j_hi], 0));
    int st = launch ? (*launch)(r) : launch_plan(c, dp, d_in, d_out, &r);
    if (st) return bail(st);
    MB_PIPE(cudaEventRecord(c->pipe_events[16 + k], c->stream));
    MB_PIPE(cudaStreamWaitEvent(c->out_stream, c->pipe_events[16 + k], 0));
    MB_PIPE(cudaMemcpyAsync((char *)host_out + r.o0 * out_plane, (const char *)d_out + r.o0 * out_plane,
                               (size_t)(r.o1 - r.o0) * out_plane, cudaMemcpyDeviceToHost, c->out_stream));
  }
#undef MB_PIPE
  (void)ob;
  MB_CUDA(c, cudaStreamSynchronize(c->out_stream));
  MB_CUDA(c, cudaStreamSynchronize(c->stream));
  return MB_OK;
}

}  // namespace mb

using namespace mb;

extern "C" {

int mb_version(void) { return MB_VERSION; }

const char *mb_status_string(int s) {
  switch (s) {
    case MB_OK: return "ok";
    case MB_ERR_INVALID_DESC: return "invalid descriptor";
    case MB_ERR_SIZE_MISMATCH: return "size mismatch";
    case MB_ERR_INDEX_ZERO: return "index zero (empty dimension under a non-Fill border)";
    case MB_ERR_UNSUPPORTED: return "unsupported element type / stencil kind";
    case MB_ERR_CUDA: return "CUDA error";
    case MB_ERR_NCCL: return "NCCL error";
    case MB_ERR_OOM: return "out of device memory";
  }
  return "unknown status";
}

int mb_elem_size(int elem) {
  switch (elem) {
    case MB_U8: case MB_I8: return 1;
    case MB_U16: case MB_I16: return 2;
    case MB_U32: case MB_I32: case MB_F32: case MB_BOOL32: return 4;
    case MB_U64: case MB_I64: case MB_F64: return 8;
  }
  return 0;
}

int mb_out_dims(const mb_stencil_desc *desc, const int64_t *in_dims, int64_t *out_dims) {
  if (!desc || !in_dims || !out_dims) return MB_ERR_INVALID_DESC;
  Plan p;
  int st = plan_geometry(desc, in_dims, &p, nullptr);
  if (st) return st;
  for (int i = 0; i < desc->rank; ++i) out_dims[i] = p.out_dims[i];
  return MB_OK;
}

int mb_border_index(int border, int64_t extent, int64_t i, int64_t *repaired, int32_t *is_fill) {
  if (!repaired || !is_fill) return MB_ERR_INVALID_DESC;
  return border_index(border, extent, i, repaired, is_fill);
}

int mb_ctx_create(int device, mb_ctx **out) {
  if (!out) return MB_ERR_INVALID_DESC;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) return MB_ERR_CUDA;  // no CPU fallback
  if (cudaSetDevice(device) != cudaSuccess) return MB_ERR_CUDA;
  mb_ctx *h = new (std::nothrow) mb_ctx();
  if (!h) return MB_ERR_OOM;
  Ctx *c = &h->c;
  c->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->num_sms = prop.multiProcessorCount;
  // (measured: giving the halo-exchange stream a higher priority is WORSE — NCCL's send/recv kernel
  // then occupies SM slots while it waits for the peer and the persistent interior kernel starts late)
  if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&c->ev_start) != cudaSuccess || cudaEventCreate(&c->ev_stop) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_a, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_b, cudaEventDisableTiming) != cudaSuccess) {
    delete h;
    return MB_ERR_CUDA;
  }
  c->stream = c->own_stream;
  if (const char *fg = std::getenv("MASSIV_B200_FORCE_GENERIC")) c->force_generic = std::atoi(fg);
  if (const char *nt = std::getenv("MASSIV_B200_NO_TMA")) c->no_tma = std::atoi(nt);
  if (const char *tb = std::getenv("MASSIV_B200_TB2")) c->tb2 = std::atoi(tb);
  *out = h;
  return MB_OK;
}

int mb_ctx_destroy(mb_ctx *h) {
  if (!h) return MB_OK;
  Ctx *c = &h->c;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  shard_destroy(c);
  c->plans.clear();
  for (int i = 0; i < 3; ++i) if (c->scratch[i]) cudaFree(c->scratch[i]);
  if (c->pinned) cudaFreeHost(c->pinned);
  if (c->life_bound) cudaFree(c->life_bound);
  if (c->life_flags) cudaFree(c->life_flags);
  if (c->ev_start) cudaEventDestroy(c->ev_start);
  if (c->ev_stop) cudaEventDestroy(c->ev_stop);
  if (c->ev_a) cudaEventDestroy(c->ev_a);
  if (c->ev_b) cudaEventDestroy(c->ev_b);
  if (c->d_flag) cudaFree(c->d_flag);
  if (c->d_conv) cudaFree(c->d_conv);
  for (cudaEvent_t e : c->pipe_events) cudaEventDestroy(e);
  if (c->out_stream) cudaStreamDestroy(c->out_stream);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  delete h;
  return MB_OK;
}

const char *mb_last_error(mb_ctx *h) { return h ? h->c.err.c_str() : "null context"; }

int mb_ctx_set_stream(mb_ctx *h, void *s) {
  Guard g(h);
  if (g.st) return g.st;
  // work queued on the old stream (scratch reuse, frees) is not ordered against the new one: drain it
  MB_CUDA(g.c, cudaStreamSynchronize(g.c->stream));
  g.c->stream = s ? (cudaStream_t)s : g.c->own_stream;
  return MB_OK;
}

int mb_ctx_set_option(mb_ctx *h, const char *key, int64_t value) {
  Guard g(h);
  if (g.st) return g.st;
  std::string k = key ? key : "";
  if (k == "force_generic") g.c->force_generic = (int)value;
  else if (k == "life_bitpack") g.c->life_bitpack = (int)value;
  else if (k == "tile_min_elems") g.c->tile_min_elems = value;
  else if (k == "pipeline") g.c->pipeline = (int)value;
  else if (k == "no_optimistic") g.c->no_optimistic = (int)value;
  else if (k == "no_tma") g.c->no_tma = (int)value;
  else if (k == "no_known") g.c->no_known = (int)value;
  else if (k == "tb2") g.c->tb2 = (int)value;
  else if (k == "no_p2p") g.c->no_p2p = (int)value;
  else if (k == "prefer_smallwin") g.c->prefer_smallwin = (int)value;
  else if (k == "no_pitch") g.c->no_pitch = (int)value;
  else if (k == "wavefront_min_mb") g.c->wavefront_min_bytes = (size_t)value << 20;
  else return fail(g.c, MB_ERR_INVALID_DESC, "unknown option " + k);
  return MB_OK;
}

// ---- the schedule of the wavefront (host arithmetic only; mb_wavefront_* export it for the CPU tests) ----
// planes [0, wf_arrived(D, nb, k)) of the input have arrived once block k is up
static int64_t wf_arrived(int64_t D, int nb, int k) { return k < 0 ? 0 : D * (int64_t)(k + 1) / nb; }
// output planes [*o0, *o1) of pass q (0-based; every pass reads two planes either side) that become computable with block k
static void wf_region(int64_t D, int nb, int k, int64_t q, int64_t *o0, int64_t *o1) {
  const int64_t back = 2 * (q + 1);
  *o0 = k == 0 ? 0 : std::max<int64_t>(0, wf_arrived(D, nb, k - 1) - back);
  *o1 = k == nb - 1 ? D : std::max<int64_t>(0, wf_arrived(D, nb, k) - back);
  if (*o1 < *o0) *o1 = *o0;
}
// planes [0, wf_final(...)) hold the result of the last of npass passes after block k's passes have run
static int64_t wf_final(int64_t D, int nb, int k, int64_t npass) {
  return k == nb - 1 ? D : std::max<int64_t>(0, wf_arrived(D, nb, k) - 2 * npass);
}

int64_t mb_wavefront_arrived(int64_t planes, int n_blocks, int block) { return wf_arrived(planes, n_blocks, block); }
void mb_wavefront_region(int64_t planes, int n_blocks, int block, int64_t pass, int64_t *o0, int64_t *o1) {
  wf_region(planes, n_blocks, block, pass, o0, o1);
}
int64_t mb_wavefront_final(int64_t planes, int n_blocks, int block, int64_t n_passes) { return wf_final(planes, n_blocks, block, n_passes); }

int mb_ctx_trim(mb_ctx *h) {
  Guard g(h);
  if (g.st) return g.st;
  Ctx *c = g.c;
  MB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->comm_stream) MB_CUDA(c, cudaStreamSynchronize(c->comm_stream));
  if (c->out_stream) MB_CUDA(c, cudaStreamSynchronize(c->out_stream));
  for (int i = 0; i < (int)(sizeof(c->scratch) / sizeof(c->scratch[0])); ++i) {
    if (c->scratch[i]) MB_CUDA(c, cudaFree(c->scratch[i]));
    c->scratch[i] = nullptr;
    c->scratch_bytes[i] = 0;
  }
  return MB_OK;
}

int mb_sync(mb_ctx *h) {
  Guard g(h);
  if (g.st) return g.st;
  MB_CUDA(g.c, cudaStreamSynchronize(g.c->stream));
  MB_CUDA(g.c, cudaStreamSynchronize(g.c->comm_stream));
  return MB_OK;
}

int64_t mb_launch_count(mb_ctx *h) { return h ? h->c.launches : 0; }
int64_t mb_aux_launch_count(mb_ctx *h) { return h ? h->c.aux_launches : 0; }
const char *mb_last_kernel(mb_ctx *h) { return h ? h->c.last_kernel : "none"; }

int mb_buf_alloc(mb_ctx *h, size_t bytes, void **p) {
  Guard g(h);
  if (g.st) return g.st;
  if (!p) return fail(g.c, MB_ERR_INVALID_DESC, "null out pointer");
  *p = nullptr;
  MB_CUDA(g.c, cudaMalloc(p, bytes ? bytes : 1));
  return MB_OK;
}

int mb_buf_free(mb_ctx *h, void *p) {
  Guard g(h);
  if (g.st) return g.st;
  if (!p) return MB_OK;
  MB_CUDA(g.c, cudaStreamSynchronize(g.c->stream));
  MB_CUDA(g.c, cudaFree(p));
  return MB_OK;
}

int mb_upload(mb_ctx *h, void *dst, const void *src, size_t bytes) {
  Guard g(h);
  if (g.st) return g.st;
  int st = copy_h2d(g.c, dst, src, bytes);
  if (st) return st;
  MB_CUDA(g.c, cudaStreamSynchronize(g.c->stream));
  return MB_OK;
}

int mb_download(mb_ctx *h, void *dst, const void *src, size_t bytes) {
  Guard g(h);
  if (g.st) return g.st;
  int st = copy_d2h(g.c, dst, src, bytes);
  if (st) return st;
  MB_CUDA(g.c, cudaStreamSynchronize(g.c->stream));
  return MB_OK;
}

// Pinned host memory is first touched (and therefore placed) by the thread that allocates it.  With eight ranks
// copying 58 GB each at once, buffers on the wrong socket share the inter-socket link: the allocation therefore
// runs with the thread bound to the CPUs of the device's NUMA node (sysfs), the old affinity is restored afterwards.
// MASSIV_B200_NO_NUMA=1 switches it off; a machine without NUMA information (numa_node = -1) is left alone.
static bool cpus_near_device(int device, cpu_set_t *set) {
  char bdf[32] = {0};
  if (cudaDeviceGetPCIBusId(bdf, (int)sizeof(bdf), device) != cudaSuccess) { cudaGetLastError(); return false; }
  for (char *q = bdf; *q; ++q) if (*q >= 'A' && *q <= 'Z') *q = (char)(*q - 'A' + 'a');
  char path[160];
  std::snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bdf);
  FILE *f = std::fopen(path, "r");
  if (!f) return false;
  int node = -1;
  const int got = std::fscanf(f, "%d", &node);
  std::fclose(f);
  if (got != 1 || node < 0) return false;
  std::snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
  f = std::fopen(path, "r");
  if (!f) return false;
  char list[1024] = {0};
  const bool ok = std::fgets(list, sizeof(list), f) != nullptr;
  std::fclose(f);
  if (!ok) return false;
  CPU_ZERO(set);
  int n = 0;
  for (char *q = list; *q;) {  // "0-31,64-95"
    char *e = nullptr;
    long a = std::strtol(q, &e, 10);
    if (e == q) break;
    long b = a;
    if (*e == '-') { q = e + 1; b = std::strtol(q, &e, 10); }
    for (long v = a; v <= b && v < CPU_SETSIZE; ++v) { CPU_SET((int)v, set); ++n; }
    q = (*e == ',') ? e + 1 : e;
    if (*e != ',' ) break;
  }
  return n > 0;
}

int mb_host_alloc(mb_ctx *h, size_t bytes, void **p) {
  Guard g(h);
  if (g.st) return g.st;
  if (!p) return fail(g.c, MB_ERR_INVALID_DESC, "null out pointer");
  cpu_set_t old_set, near_set;
  bool bound = false;
  if (!std::getenv("MASSIV_B200_NO_NUMA") && sched_getaffinity(0, sizeof(old_set), &old_set) == 0 &&
      cpus_near_device(g.c->device, &near_set)) {
    cpu_set_t both;
    CPU_AND(&both, &old_set, &near_set);  // never leave the set the process was given
    if (CPU_COUNT(&both) > 0 && sched_setaffinity(0, sizeof(both), &both) == 0) bound = true;
  }
  const cudaError_t e = cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocPortable);
  if (e == cudaSuccess && bound && bytes) {
    // first touch on this thread: one byte per page places the pages on the near node
    volatile unsigned char *q = (volatile unsigned char *)*p;
    for (size_t i = 0; i < bytes; i += 4096) q[i] = 0;
  }
  if (bound) sched_setaffinity(0, sizeof(old_set), &old_set);
  MB_CUDA(g.c, e);
  return MB_OK;
}

int mb_host_free(mb_ctx *h, void *p) {
  Guard g(h);
  if (g.st) return g.st;
  if (p) MB_CUDA(g.c, cudaFreeHost(p));
  return MB_OK;
}

int mb_run_dev(mb_ctx *h, const mb_stencil_desc *desc, const int64_t *in_dims, const void *d_in, void *d_out) {
  Guard g(h);
  if (g.st) return g.st;
  if (!desc || !in_dims) return fail(g.c, MB_ERR_INVALID_DESC, "null argument");
  std::shared_ptr<DevPlan> dp;
  int st = get_dev_plan(g.c, desc, in_dims, &dp);
  if (st) return st;
  return launch_plan(g.c, *dp, d_in, d_out, nullptr);
}

int mb_run(mb_ctx *h, const mb_stencil_desc *desc, const int64_t *in_dims, const void *host_in, void *host_out) {
  Guard g(h);
  if (g.st) return g.st;
  Ctx *c = g.c;
  if (!desc || !in_dims) return fail(c, MB_ERR_INVALID_DESC, "null argument");
  std::shared_ptr<DevPlan> dp;
  int st = get_dev_plan(c, desc, in_dims, &dp);
  if (st) return st;
  const size_t ib = (size_t)dp->p.in_elems * dp->p.esize, ob = (size_t)dp->p.out_elems * dp->p.esize;
  void *d_in = nullptr, *d_out = nullptr;
  if ((st = ensure_scratch(c, 0, ib ? ib : 1, &d_in))) return st;
  if ((st = ensure_scratch(c, 1, ob ? ob : 1, &d_out))) return st;
  if ((st = run_pipelined(c, *dp, host_in, host_out, d_in, d_out)) != -1) return st;
  if ((st = copy_h2d(c, d_in, host_in, ib))) return st;
  if ((st = launch_plan(c, *dp, d_in, d_out, nullptr))) return st;
  if ((st = copy_d2h(c, host_out, d_out, ob))) return st;
  MB_CUDA(c, cudaStreamSynchronize(c->stream));
  return MB_OK;
}

int mb_run_sep2_dev(mb_ctx *h, const mb_stencil_desc *f, const mb_stencil_desc *s, const int64_t *dims,
                    const void *d_in, void *d_out, void *d_tmp) {
  Guard g(h);
  if (g.st) return g.st;
  return sep2_dev_locked(g.c, f, s, dims, d_in, d_out, d_tmp);
}

int mb_run_sep2(mb_ctx *h, const mb_stencil_desc *f, const mb_stencil_desc *s, const int64_t *dims,
                const void *host_in, void *host_out) {
  Guard g(h);
  if (g.st) return g.st;
  Ctx *c = g.c;
  int st = check_sep2(c, f, s);
  if (st) return st;
  size_t n = (size_t)mb_elem_size(f->elem);
  for (int i = 0; i < f->rank; ++i) n *= (size_t)dims[i];
  void *d_in = nullptr, *d_out = nullptr;
  if ((st = ensure_scratch(c, 0, n ? n : 1, &d_in))) return st;
  if ((st = ensure_scratch(c, 1, n ? n : 1, &d_out))) return st;
  {
    // the fused kernel produces any range of output rows from the input rows it reads (the column pass's reach):
    // the same three-stage pipeline as mb_run, driven by the second pass's geometry
    std::shared_ptr<DevPlan> p1, p2;
    if ((st = get_dev_plan(c, f, dims, &p1))) return st;
    if ((st = get_dev_plan(c, s, dims, &p2))) return st;
    if (p1->p.same_shape() && p2->p.same_shape() && !c->force_generic && p1->p.post.empty() && p2->p.post.empty()) {
      // probe: does the fused kernel take these shapes?  (an empty range launches nothing)
      const OuterRange none{0, 0};
      if (run_sep2_fused(c, *p1, *p2, d_in, d_out, &none) == MB_OK) {
        const std::function<int(const OuterRange &)> launch = [&](const OuterRange &r) { return run_sep2_fused(c, *p1, *p2, d_in, d_out, &r); };
        if ((st = run_pipelined(c, *p2, host_in, host_out, d_in, d_out, &launch)) != -1) return st;
      }
    }
  }
  if ((st = copy_h2d(c, d_in, host_in, n))) return st;
  if ((st = sep2_dev_locked(c, f, s, dims, d_in, d_out, nullptr))) return st;
  if ((st = copy_d2h(c, host_out, d_out, n))) return st;
  MB_CUDA(c, cudaStreamSynchronize(c->stream));
  return MB_OK;
}

int mb_iterate_dev(mb_ctx *h, const mb_stencil_desc *desc, const int64_t *dims, void *a, void *b,
                   int64_t n_steps, void **result) {
  Guard g(h);
  if (g.st) return g.st;
  if (!desc || !dims || !result) return fail(g.c, MB_ERR_INVALID_DESC, "null argument");
  return iterate_dev_locked(g.c, desc, dims, a, b, n_steps, result);
}

// Host -> host iteration of the two-step kernel's plans with the copies hidden behind the passes.
// A pass over planes [lo, hi) reads the previous pass's planes [lo - 2, hi + 2), so after the first Z planes of the
// input have arrived pass p can already produce the planes below Z - 2 (p + 1): the dependency cone of a stencil.  The
// input travels up in NB blocks; after block k (planes below Z_k) has landed, EVERY pass runs once, pass p over
// [Z_(k-1) - 2 (p + 1), Z_k - 2 (p + 1)) — a parallelogram in the (plane, pass) diagram, one launch of about D / NB
// planes per pass and block, all on the compute stream in that order (which is also a valid order for the two
// ping-pong arrays: a plane only ever moves up in pass level).  Block k + 1 travels while block k's passes run, and the
// planes whose last pass is done leave for the host while the later blocks' passes run.  What stays exposed is the
// first block's upload and the last 2 P + D / NB planes' download.
// No finite promise: one flag collects the kernels' reports over the whole job; if it is set at the end the job is
// redone the plain way (upload, checked passes with their strict redo, download).
// -1: not applicable (the caller runs the plain way).
static int iterate_wavefront(Ctx *c, const DevPlan *dp, const void *host_in, void *host_out, void *a, void *b, int64_t n_steps,
                             int64_t pitch) {
  const Plan &p = dp->p;
  if (!c->pipeline || c->force_generic || p.rank != 3 || !p.same_shape() || !p.post.empty()) return -1;
  const int tb2 = vol3d_tb2_mode(c, p);
  const int64_t D = p.in_dims[0], pairs = n_steps / 2, npass = pairs + (n_steps & 1);
  const int64_t row_elems = pitch ? pitch : p.in_dims[2];
  const size_t plane_dev = (size_t)p.in_dims[1] * (size_t)row_elems * p.esize;       // bytes per plane on the device
  const size_t plane_host = (size_t)p.in_dims[1] * (size_t)p.in_dims[2] * p.esize;  // ... in the caller's arrays
  if (!tb2 || pairs < 2 || D < 16 || plane_host * (size_t)D < c->wavefront_min_bytes) return -1;
  if (tb2 == 2 && ensure_flag(c) != MB_OK) return -1;
  int NB = 8;  // (measured: 4 / 8 / 16 blocks — see DESIGN.md; MASSIV_B200_WF_BLOCKS is the measurement knob)
  if (const char *e = std::getenv("MASSIV_B200_WF_BLOCKS")) NB = std::max(2, std::min(16, std::atoi(e)));
  if (!c->out_stream && cudaStreamCreateWithFlags(&c->out_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return -1; }
  while ((int)c->pipe_events.size() < 2 * 16) {
    cudaEvent_t e;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return -1; }
    c->pipe_events.push_back(e);
  }
  unsigned *flag = tb2 == 2 ? c->d_flag + 16 : nullptr;  // (its own word: single-step launches reset c->d_flag[0])
  auto bail = [&](int st) {
    cudaStreamSynchronize(c->comm_stream);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->out_stream);
    cudaGetLastError();
    c->in_pitch = c->out_pitch = 0;
    return st;
  };
#define MB_WF(expr)                                        \
  do {                                                     \
    int _ps = ::mb::check_cuda(c, (expr), #expr);          \
    if (_ps) return bail(_ps);                             \
  } while (0)
  MB_WF(cudaEventRecord(c->ev_a, c->stream));  // whatever the caller queued on the compute stream comes first
  MB_WF(cudaStreamWaitEvent(c->comm_stream, c->ev_a, 0));
  MB_WF(cudaStreamWaitEvent(c->out_stream, c->ev_a, 0));
  if (flag) MB_WF(cudaMemsetAsync(flag, 0, sizeof(unsigned), c->stream));
  const size_t rb = (size_t)p.in_dims[2] * p.esize, pb = (size_t)row_elems * p.esize;
  auto Z = [&](int k) { return wf_arrived(D, NB, k); };  // planes [0, Z(k)) have arrived after block k
  // uploads, in order, on the copy stream
  for (int k = 0; k < NB; ++k) {
    const int64_t z0 = k ? Z(k - 1) : 0, z1 = Z(k);
    const char *src = (const char *)host_in + (size_t)z0 * plane_host;
    char *dst = (char *)a + (size_t)z0 * plane_dev;
    if (pitch) MB_WF(cudaMemcpy2DAsync(dst, pb, src, rb, rb, (size_t)(z1 - z0) * (size_t)p.in_dims[1], cudaMemcpyHostToDevice, c->comm_stream));
    else MB_WF(cudaMemcpyAsync(dst, src, (size_t)(z1 - z0) * plane_host, cudaMemcpyHostToDevice, c->comm_stream));
    MB_WF(cudaEventRecord(c->pipe_events[k], c->comm_stream));
  }
  void *fin = (npass & 1) ? b : a;  // where the last pass writes
  c->in_pitch = c->out_pitch = pitch;
  int64_t sent = 0;  // planes [0, sent) of the result are on their way to the host
  for (int k = 0; k < NB; ++k) {
    MB_WF(cudaStreamWaitEvent(c->stream, c->pipe_events[k], 0));
    for (int64_t q = 0; q < npass; ++q) {
      OuterRange r;
      wf_region(D, NB, k, q, &r.o0, &r.o1);
      if (r.o1 <= r.o0) continue;
      void *in = (q & 1) ? b : a, *out = (q & 1) ? a : b;
      int st = q < pairs ? launch_vol3d_tb2(c, p, in, out, &r, 0, 0, D, flag) : launch_plan(c, *dp, in, out, &r);
      if (st == -1) st = fail(c, MB_ERR_UNSUPPORTED, "internal: two-step kernel declined a block of the wavefront");
      if (st) return bail(st);
    }
    // the planes whose last pass is done leave while the next blocks' passes run
    const int64_t done = wf_final(D, NB, k, npass);
    if (done > sent) {
      MB_WF(cudaEventRecord(c->pipe_events[16 + k], c->stream));
      MB_WF(cudaStreamWaitEvent(c->out_stream, c->pipe_events[16 + k], 0));
      char *dst = (char *)host_out + (size_t)sent * plane_host;
      const char *src = (const char *)fin + (size_t)sent * plane_dev;
      if (pitch) MB_WF(cudaMemcpy2DAsync(dst, rb, src, pb, rb, (size_t)(done - sent) * (size_t)p.in_dims[1], cudaMemcpyDeviceToHost, c->out_stream));
      else MB_WF(cudaMemcpyAsync(dst, src, (size_t)(done - sent) * plane_host, cudaMemcpyDeviceToHost, c->out_stream));
      sent = done;
    }
  }
  c->in_pitch = c->out_pitch = 0;
  c->last_kernel = "vol3d_tb2";
  // join: later work on the compute stream (and the caller) sees everything done
  MB_WF(cudaEventRecord(c->ev_b, c->out_stream));
  MB_WF(cudaStreamWaitEvent(c->stream, c->ev_b, 0));
  unsigned seen = 0;
  if (flag) MB_WF(cudaMemcpyAsync(&seen, flag, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
  MB_WF(cudaStreamSynchronize(c->stream));
#undef MB_WF
  return seen ? -2 : MB_OK;  // -2: Inf / NaN met on the way — redo strictly
}

int mb_iterate(mb_ctx *h, const mb_stencil_desc *desc, const int64_t *dims, const void *host_in, void *host_out,
               int64_t n_steps) {
  Guard g(h);
  if (g.st) return g.st;
  Ctx *c = g.c;
  if (!desc || !dims) return fail(c, MB_ERR_INVALID_DESC, "null argument");
  size_t n = (size_t)mb_elem_size(desc->elem);
  for (int i = 0; i < desc->rank; ++i) n *= (size_t)dims[i];
  void *a = nullptr, *b = nullptr, *res = nullptr;
  int st;
  // the device layout is ours to choose: rows padded to 16-byte units where that lets the iteration use TMA
  // (iterate_pitch); the copies place and gather the rows
  if (n_steps >= 1 && n) {
    std::shared_ptr<DevPlan> dp;
    if ((st = get_dev_plan(c, desc, dims, &dp))) return st;
    const int64_t pitch = iterate_pitch(c, dp.get(), n_steps);
    if (pitch) {
      const Plan &p = dp->p;
      const size_t rows = (size_t)(p.in_dims[0] * p.in_dims[1]), rb = (size_t)p.in_dims[2] * p.esize, pb = (size_t)pitch * p.esize;
      if ((st = ensure_scratch(c, 0, rows * pb, &a))) return st;
      if ((st = ensure_scratch(c, 1, rows * pb, &b))) return st;
      if ((st = iterate_wavefront(c, dp.get(), host_in, host_out, a, b, n_steps, pitch)) >= 0) return st;
      if ((st = copy_rows(c, a, pb, host_in, rb, rb, rows, cudaMemcpyHostToDevice))) return st;
      c->in_pitch = c->out_pitch = pitch;
      st = iterate_core(c, dp.get(), a, b, n_steps, &res);
      c->in_pitch = c->out_pitch = 0;
      if (st) return st;
      if ((st = copy_rows(c, host_out, rb, res, pb, rb, rows, cudaMemcpyDeviceToHost))) return st;
      MB_CUDA(c, cudaStreamSynchronize(c->stream));
      return MB_OK;
    }
  }
  if ((st = ensure_scratch(c, 0, n ? n : 1, &a))) return st;
  if ((st = ensure_scratch(c, 1, n ? n : 1, &b))) return st;
  if (n_steps >= 1 && n) {
    std::shared_ptr<DevPlan> dp;
    if ((st = get_dev_plan(c, desc, dims, &dp))) return st;
    if (dp->p.same_shape() && (st = iterate_wavefront(c, dp.get(), host_in, host_out, a, b, n_steps, 0)) >= 0) return st;
  }
  if ((st = copy_h2d(c, a, host_in, n))) return st;
  if ((st = iterate_dev_locked(c, desc, dims, a, b, n_steps, &res))) return st;
  if ((st = copy_d2h(c, host_out, res, n))) return st;
  MB_CUDA(c, cudaStreamSynchronize(c->stream));
  return MB_OK;
}

int mb_iterate_until_dev(mb_ctx *h, const mb_stencil_desc *desc, const int64_t *dims, void *a, void *b, int64_t max_steps,
                         int64_t check_every, int predicate, double tol, int64_t *steps, int32_t *converged, void **result) {
  Guard g(h);
  if (g.st) return g.st;
  if (!desc || !dims || !steps || !converged || !result) return fail(g.c, MB_ERR_INVALID_DESC, "null argument");
  return iterate_until_locked(g.c, desc, dims, a, b, max_steps, check_every, predicate, tol, steps, converged, result);
}

int mb_iterate_until(mb_ctx *h, const mb_stencil_desc *desc, const int64_t *dims, const void *host_in, void *host_out,
                     int64_t max_steps, int64_t check_every, int predicate, double tol, int64_t *steps, int32_t *converged) {
  Guard g(h);
  if (g.st) return g.st;
  Ctx *c = g.c;
  if (!desc || !dims || !steps || !converged) return fail(c, MB_ERR_INVALID_DESC, "null argument");
  size_t n = (size_t)mb_elem_size(desc->elem);
  for (int i = 0; i < desc->rank; ++i) n *= (size_t)dims[i];
  void *a = nullptr, *b = nullptr, *res = nullptr;
  int st;
  if ((st = ensure_scratch(c, 0, n ? n : 1, &a))) return st;
  if ((st = ensure_scratch(c, 1, n ? n : 1, &b))) return st;
  if ((st = copy_h2d(c, a, host_in, n))) return st;
  if ((st = iterate_until_locked(c, desc, dims, a, b, max_steps, check_every, predicate, tol, steps, converged, &res))) return st;
  if ((st = copy_d2h(c, host_out, res, n))) return st;
  MB_CUDA(c, cudaStreamSynchronize(c->stream));
  return MB_OK;
}

int mb_timer_start(mb_ctx *h) {
  Guard g(h);
  if (g.st) return g.st;
  MB_CUDA(g.c, cudaEventRecord(g.c->ev_start, g.c->stream));
  return MB_OK;
}

int mb_timer_stop(mb_ctx *h, float *ms) {
  Guard g(h);
  if (g.st) return g.st;
  MB_CUDA(g.c, cudaEventRecord(g.c->ev_stop, g.c->stream));
  MB_CUDA(g.c, cudaEventSynchronize(g.c->ev_stop));
  if (ms) MB_CUDA(g.c, cudaEventElapsedTime(ms, g.c->ev_start, g.c->ev_stop));
  return MB_OK;
}

}  // extern "C"
