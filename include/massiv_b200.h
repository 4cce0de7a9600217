/*
 * massiv_b200.h — C ABI of libmassiv_b200.so: massiv's stencil / windowed-load hot path on B200.
 *
 * The reference (lehins/massiv 1.0.5.0) is pure Haskell with no FFI on this path; the seam this
 * library replaces is `compute` of an `Array DW` produced by mapStencil/applyStencil, i.e. the
 * per-element callback loader
 *     iterArrayLinearST_ :: Scheduler s () -> Array DW ix e -> (Int -> e -> ST s ()) -> ST s ()
 *     (massiv/src/Data/Massiv/Core/Common.hs:420-426; instances Array/Delayed/Windowed.hs:231,350,364).
 * A Haskell shim (haskell/Data/Massiv/Array/Stencil/B200.hs, bindings shown in INTEGRATION.md)
 * reifies the built-in stencil constructors into `mb_stencil_desc` and calls the entry points
 * below with `foreign import ccall safe` on Storable (`S`) arrays.  Paths in the comments are
 * relative to /root/reference/massiv/src/Data/Massiv/.
 *
 * Conventions
 *   - arrays are dense row-major, last index fastest, no pitch (Core/Index/Internal.hs:584-595);
 *     dims[0] is the OUTERMOST dimension; native-endian elements; element sizes follow
 *     Storable.sizeOf (Bool is 4 bytes, Array/Manifest/Storable.hs:71-77).
 *   - every function returns an mb_status; no C++ exception crosses the boundary;
 *     mb_last_error(ctx) gives the text.  The shim maps MB_ERR_INDEX_ZERO to IndexZeroException,
 *     MB_ERR_SIZE_MISMATCH to SizeMismatchException (Core/Index/Internal.hs:1136-1142).
 *   - re-entrant: a context serialises its own calls with a mutex and sets its device on every
 *     entry, so it may be called from any OS thread (`compute` is unsafePerformIO and safe FFI calls
 *     migrate between OS threads, Array/Manifest/Internal.hs:66).
 *   - there is NO CPU fallback: every entry point that computes fails with MB_ERR_CUDA when no
 *     device is usable.
 *   - floating point: one IEEE-754 round-to-nearest operation per reference operation, never
 *     contracted into FMA, in the reference's accumulation order: results
 *     are bit-identical to massiv's for finite and non-finite inputs (NaN payloads excepted).
 *     MB_FLAG_ASSUME_FINITE lets CONV/CORR/TAPS/MAG2 skip zero coefficients, which is
 *     bit-identical for finite inputs only (a skipped 0*Inf would have produced NaN).  The flag is
 *     never needed for correctness or speed of the star-shaped 3x3x3 kernels: without it they skip
 *     the zero taps optimistically, check on the device that no Inf / NaN was involved, and redo the
 *     work with the strict 27-tap kernel otherwise.
 */
#ifndef MASSIV_B200_H
#define MASSIV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB_MAX_RANK 5
#define MB_VERSION 100 /* 0.1.0 */

typedef enum mb_status {
  MB_OK = 0,
  MB_ERR_INVALID_DESC = 1,  /* malformed descriptor / arguments                              */
  MB_ERR_SIZE_MISMATCH = 2, /* iterate on a stencil whose output size differs from its input  */
  MB_ERR_INDEX_ZERO = 3,    /* non-Fill border evaluated on a zero extent: IndexZeroException
                               (Core/Index/Internal.hs:1031)                                   */
  MB_ERR_UNSUPPORTED = 4,   /* element type / kind combination the reference's classes forbid */
  MB_ERR_CUDA = 5,
  MB_ERR_NCCL = 6,
  MB_ERR_OOM = 7
} mb_status;

/* Storable element types */
typedef enum mb_elem {
  MB_U8 = 0, MB_I8, MB_U16, MB_I16, MB_U32, MB_I32, MB_U64, MB_I64, MB_F32, MB_F64, MB_BOOL32
} mb_elem;

/* the stencil constructors the device path recognises (closures cannot run on the device) */
typedef enum mb_kind {
  MB_ID = 0,   /* idStencil                                   Array/Stencil.hs:269            */
  MB_SUM,      /* sumStencil      ((0 + x0) + x1) + ...        Array/Stencil.hs:425-427        */
  MB_PRODUCT,  /* productStencil                               Array/Stencil.hs:454-456        */
  MB_AVG,      /* avgStencil      sum / fromIntegral n         Array/Stencil.hs:479-481        */
  MB_MAX,      /* maxStencil      from minBound                Array/Stencil.hs:368-370        */
  MB_MIN,      /* minStencil      from maxBound                Array/Stencil.hs:401-403        */
  MB_CONV,     /* makeConvolutionStencilFromKernel             Array/Stencil/Convolution.hs:64-80   */
  MB_CORR,     /* makeCorrelationStencilFromKernel             Array/Stencil/Convolution.hs:107-121 */
  MB_TAPS,     /* make(Unsafe){Convolution,Correlation}Stencil closure form lowered to an ordered
                  tap list: acc = src[ix + off_t] * k_t + acc  Array/Stencil/Convolution.hs:44-57, 89-100 */
  MB_LIFE,     /* Game-of-Life rule on Word8                   massiv-examples/GameOfLife/app/GameOfLife.hs:15-36 */
  MB_MAG2,     /* sqrt (fmap (^2) a + fmap (^2) b), a,b FromKernel stencils of one size (Sobel)
                  Array/Stencil/Internal.hs:105-146; massiv-bench/src/Data/Massiv/Bench/Sobel.hs:39-44 */
  /* Integral approximation rules (Array/Numeric/Integral.hs): 1-D stencils along one dimension
   * (size is 1 elsewhere), centre 0, coeffs[0] = dx.  g_i is the i-th sample along that dimension. */
  MB_INT_MIDPOINT,  /* size k:    dx * (((0 + g_0) + g_1) + ... + g_{k-1})              Integral.hs:54-66   */
  MB_INT_TRAPEZOID, /* size n+1:  (dx/2) * ((g_0 + 2*g_1 + ... + 2*g_{n-1}) + g_n)       Integral.hs:75-90   */
  MB_INT_SIMPSON,   /* size n+1, n even >= 2: (dx/3) * acc, acc = ((acc + g_i) + 4*g_{i+1}) + g_{i+2}
                       over i = 0, 2, .., n-2                                           Integral.hs:99-118  */
  /* A stencil EXPRESSION: the Applicative / Num / Fractional / Floating instances of Stencil
   * (Array/Stencil/Internal.hs:83-174).  `terms` are sub-stencils of the kinds above (ID .. TAPS), every one
   * evaluated at the same absolute source index; `program` combines their values point-wise in postfix order
   * (liftA2 (+|-|*|/), fmap f, pure c).  size / center of the descriptor are the union the reference computes:
   * centre = max of the terms' centres, size = centre + max (size_t - centre_t) (unionStencilCenters / Sizes,
   * Internal.hs:83-90); they are validated.  Example: the closure-form Sobel operator of the reference's own
   * benchmark, sqrt (sX * sX + sY * sY) (massiv-bench/src/Data/Massiv/Bench/Sobel.hs:15-44), is two TAPS terms
   * and the program  T0 T0 MUL T1 T1 MUL ADD SQRT. */
  MB_EXPR
} mb_kind;

/* Border (Core/Index.hs:159-195) resolved by handleBorderIndex (Core/Index.hs:236-253) */
typedef enum mb_border { MB_FILL = 0, MB_WRAP, MB_EDGE, MB_REFLECT, MB_CONTINUE } mb_border;

/* Post-maps: `fmap f` over the stencil's result (Functor (Stencil ix e), Stencil/Internal.hs:36-41,
 * 70-83) or over the windowed array (Functor (Array DW ix), Delayed/Windowed.hs:78-84), and the
 * unary methods of the Num / Fractional / Floating instances of Stencil that are defined as fmap
 * (Stencil/Internal.hs:114-146), for the f whose result is fixed by IEEE-754 / two's complement —
 * so that it stays bit-identical.  Applied in order to every output cell.  c is `operand`. */
typedef enum mb_post {
  MB_POST_NEGATE = 1, /* negate                                                        */
  MB_POST_ABS,        /* abs      (minBound stays minBound for integers)               */
  MB_POST_SIGNUM,     /* signum   (+-0 and NaN map to themselves for Float/Double)     */
  MB_POST_SQUARE,     /* (^2)     x * x                                                */
  MB_POST_ADD,        /* (+ c)    x + c                                                */
  MB_POST_SUB,        /* subtract c: x - c                                             */
  MB_POST_RSUB,       /* (c -)    c - x                                                */
  MB_POST_MUL,        /* (* c)    x * c                                                */
  MB_POST_MIN,        /* min x c = if x <= c then x else c     (GHC.Classes default)   */
  MB_POST_MAX,        /* max x c = if x <= c then c else x                             */
  MB_POST_DIV,        /* (/ c)    x / c      Float / Double only from here on          */
  MB_POST_RDIV,       /* (c /)    c / x                                                */
  MB_POST_RECIP,      /* recip    1 / x                                                */
  MB_POST_SQRT        /* sqrt                                                          */
} mb_post;
#define MB_MAX_POST 8
typedef struct mb_post_op {
  int32_t op;         /* mb_post */
  int32_t reserved;
  uint8_t operand[8]; /* c in the element type (native-endian, first sizeof(elem) bytes) */
} mb_post_op;

/* One sub-stencil of an MB_EXPR descriptor: the fields of mb_stencil_desc that define WHICH values are combined
 * (kind, size, centre, kernel / tap list); element type, rank, border, padding and stride are the expression's. */
typedef struct mb_term {
  int32_t kind;                  /* mb_kind: MB_ID, MB_SUM, MB_PRODUCT, MB_AVG, MB_MAX, MB_MIN, MB_CONV, MB_CORR, MB_TAPS */
  int32_t reserved;
  int64_t size[MB_MAX_RANK];
  int64_t center[MB_MAX_RANK];   /* stencilCenter of the term on its own (validated like a descriptor's)      */
  const void *coeffs;            /* CONV/CORR: prod(size) elements; TAPS: ntaps                               */
  int64_t ntaps;                 /* TAPS                                                                      */
  const int64_t *tap_offsets;    /* TAPS: ntaps*rank                                                          */
} mb_term;
/* Postfix program over a value stack (depth <= MB_MAX_STACK).  Binary operators pop b then a and push a OP b. */
typedef enum mb_xop {
  MB_X_TERM = 1,  /* push the value of terms[arg]                                         */
  MB_X_CONST,     /* push `operand`: pure c / fromInteger / fromRational                  */
  MB_X_ADD,       /* liftA2 (+)      Internal.hs:115                                      */
  MB_X_SUB,       /* liftA2 (-)                                                           */
  MB_X_MUL,       /* liftA2 (*)                                                           */
  MB_X_DIV,       /* liftA2 (/)      Fractional e                 Internal.hs:131         */
  MB_X_MIN,       /* liftA2 min: min a b = if a <= b then a else b (GHC.Classes default)  */
  MB_X_MAX,       /* liftA2 max: max a b = if a <= b then b else a                        */
  MB_X_NEGATE,    /* fmap negate; the unary operators follow mb_post                      */
  MB_X_ABS,
  MB_X_SIGNUM,
  MB_X_SQUARE,    /* fmap (^2): x * x                                                     */
  MB_X_RECIP,     /* fmap recip: 1 / x        Fractional e                                */
  MB_X_SQRT       /* fmap sqrt                Floating e                                  */
} mb_xop;
typedef struct mb_xins {
  int32_t op;         /* mb_xop */
  int32_t arg;        /* MB_X_TERM: index into terms */
  uint8_t operand[8]; /* MB_X_CONST: c in the element type */
} mb_xins;
#define MB_MAX_TERMS 8
#define MB_MAX_PROGRAM 32
#define MB_MAX_STACK 8

#define MB_FLAG_MAG2_CORR 1u     /* MAG2 sub-stencils are correlations                         */
#define MB_FLAG_ASSUME_FINITE 2u /* inputs are finite: zero coefficients may be skipped        */

/*
 * One reified  applyStencil (Padding pad_lo pad_hi border) stencil  [computeWithStride stride].
 * mapStencil = pad_lo := center, pad_hi := size - (center + 1)   (Array/Stencil.hs:76-86, 173-179).
 * `center` is stencilCenter: 0 for the fold stencils, size-1-(size `quot` 2) for CONV/MAG2,
 * size `div` 2 for CORR, (1,1) for LIFE, caller-chosen for TAPS; it is validated.
 */
typedef struct mb_stencil_desc {
  int32_t kind;                  /* mb_kind   */
  int32_t elem;                  /* mb_elem   */
  int32_t rank;                  /* 1..MB_MAX_RANK */
  int32_t border;                /* mb_border */
  int64_t size[MB_MAX_RANK];     /* stencilSize, outermost first                               */
  int64_t center[MB_MAX_RANK];   /* stencilCenter                                              */
  int64_t pad_lo[MB_MAX_RANK];   /* paddingFromOrigin                                          */
  int64_t pad_hi[MB_MAX_RANK];   /* paddingFromBottom                                          */
  int64_t stride[MB_MAX_RANK];   /* Stride (Core/Index/Stride.hs:52-61); 1 = plain compute     */
  uint8_t fill[8];               /* Fill element (native-endian, first sizeof(elem) bytes)     */
  const void *coeffs;            /* CONV/CORR/MAG2: prod(size) elements row-major; TAPS: ntaps;
                                    INT_*: one element, dx                                     */
  const void *coeffs2;           /* MAG2: second kernel                                        */
  int64_t ntaps;                 /* TAPS                                                       */
  const int64_t *tap_offsets;    /* TAPS: ntaps*rank source offsets in application order       */
  uint32_t flags;
  uint32_t n_post;               /* number of post-maps, <= MB_MAX_POST                       */
  const mb_post_op *post;        /* applied in order to every result                          */
  /* MB_EXPR only (zero / NULL otherwise) */
  uint32_t n_terms;              /* <= MB_MAX_TERMS                                            */
  uint32_t n_program;            /* <= MB_MAX_PROGRAM                                          */
  const mb_term *terms;
  const mb_xins *program;
} mb_stencil_desc;

typedef struct mb_ctx mb_ctx;

/* ---- library / context -------------------------------------------------------------------- */
int mb_version(void);
const char *mb_status_string(int status);
/* One context = one device + its streams, scratch and (optionally) one NCCL communicator.
 * Replaces the scheduler hand-off of withMassivScheduler_ (Core/Loop.hs:468-474). */
int mb_ctx_create(int device, mb_ctx **out);
int mb_ctx_destroy(mb_ctx *ctx);
const char *mb_last_error(mb_ctx *ctx);
/* launch the context's kernels on an external cudaStream_t (NULL restores its own stream) */
int mb_ctx_set_stream(mb_ctx *ctx, void *cuda_stream);
int mb_sync(mb_ctx *ctx);
/* Returns the context's grow-only device scratch to the driver (the staging arrays of the host -> host entry points, the
 * padded pair of an iteration over an array TMA cannot describe — up to four arrays of the largest size seen).  Waits
 * for the context's streams first.  The next call that needs scratch allocates it again. */
int mb_ctx_trim(mb_ctx *ctx);
/* tuning / test switches (none changes a result): "force_generic" (1 = always the generic kernel),
 * "life_bitpack" (0 = byte-wise Life kernels only), "tile_min_elems" (arrays below this many output
 * cells take the generic kernel), "no_tma" (stage tiles with cp.async), "no_known" (no compile-time
 * coefficient kernels), "tb2" (0 = one step per pass when iterating), "no_optimistic" (1 = never skip
 * zero taps without the caller's promise), "pipeline" (0 = serial copies in mb_run / mb_run_sep2 / mb_iterate),
 * "no_p2p" (1 = sharded iteration moves halo planes with ncclSend/ncclRecv; must be set alike on every rank),
 * "prefer_smallwin" (1 = the table-driven small-window kernel before the compile-time-shaped ones), "no_pitch"
 * (1 = iterations over rank-3 arrays without a tensor map stay in the caller's layout instead of running on an
 * internal pair with padded rows), "wavefront_min_mb" (mb_iterate hides its copies behind the passes for arrays
 * of at least this many MiB; default 256) */
int mb_ctx_set_option(mb_ctx *ctx, const char *key, int64_t value);
/* number of kernels of THIS library launched through the context so far */
int64_t mb_launch_count(mb_ctx *ctx);
/* of those, the helper launches: conditional follow-ups of the checked kernels (they return at once
 * unless a check failed), post-map passes, constant ghost-plane fills */
int64_t mb_aux_launch_count(mb_ctx *ctx);
/* name of the kernel family the last launch used ("generic", "tile2d", ...); for tests/bench */
const char *mb_last_kernel(mb_ctx *ctx);

/* ---- the block schedule of mb_iterate's hidden copies (host arithmetic only; exported for tests) ----
 * The input of `planes` outermost planes travels up in n_blocks blocks; after block k, planes
 * [0, mb_wavefront_arrived(.., k)) are on the device.  With block k every pass q (0-based; a pass reads two planes either
 * side of what it writes) produces output planes [*o0, *o1) (mb_wavefront_region), in the order block-major, pass-minor;
 * afterwards planes [0, mb_wavefront_final(.., k, n_passes)) hold the final result and may leave. */
int64_t mb_wavefront_arrived(int64_t planes, int n_blocks, int block);
void mb_wavefront_region(int64_t planes, int n_blocks, int block, int64_t pass, int64_t *o0, int64_t *o1);
int64_t mb_wavefront_final(int64_t planes, int n_blocks, int block, int64_t n_passes);

/* ---- geometry (host only; usable without a GPU) -------------------------------------------- */
int mb_elem_size(int elem);
/* size arithmetic of applyStencil (Array/Stencil.hs:207-208) and strideSize (Stride.hs:102-105) */
int mb_out_dims(const mb_stencil_desc *desc, const int64_t *in_dims, int64_t *out_dims);
/* handleBorderIndex for one dimension; *is_fill = 1 when Fill yields the element instead */
int mb_border_index(int border, int64_t extent, int64_t i, int64_t *repaired, int32_t *is_fill);

/* ---- memory --------------------------------------------------------------------------------- */
int mb_buf_alloc(mb_ctx *ctx, size_t bytes, void **dev_ptr);
int mb_buf_free(mb_ctx *ctx, void *dev_ptr);
int mb_upload(mb_ctx *ctx, void *dev_dst, const void *host_src, size_t bytes);
int mb_download(mb_ctx *ctx, void *host_dst, const void *dev_src, size_t bytes);
/* pinned host memory an `S` array can wrap (unsafeArrayFromForeignPtr0, Storable.hs:342-344) */
int mb_host_alloc(mb_ctx *ctx, size_t bytes, void **host_ptr);
int mb_host_free(mb_ctx *ctx, void *host_ptr);

/* ---- compute: the replacement of `compute (applyStencil ...)` ------------------------------ */
/* host -> host, the literal drop-in of computeAs S . applyStencil (Manifest/Internal.hs:65-126).
 * host_out must hold prod(mb_out_dims) elements. */
int mb_run(mb_ctx *ctx, const mb_stencil_desc *desc, const int64_t *in_dims, const void *host_in,
           void *host_out);
/* device-resident form (asynchronous on the context stream) */
int mb_run_dev(mb_ctx *ctx, const mb_stencil_desc *desc, const int64_t *in_dims, const void *dev_in,
               void *dev_out);
/* two-pass separable application: compute (mapStencil b second) . compute (mapStencil b first),
 * the intermediate rounded to the element type — the massiv idiom for rank-1 factorable kernels.
 * `first`/`second` must be same-size (mapStencil-shaped) descriptors of one elem/rank/border.
 * Fused into one HBM round trip when the shapes allow, otherwise run as two passes through
 * dev_tmp (may be NULL: scratch is taken from the context). */
int mb_run_sep2_dev(mb_ctx *ctx, const mb_stencil_desc *first, const mb_stencil_desc *second,
                    const int64_t *dims, const void *dev_in, void *dev_out, void *dev_tmp);
int mb_run_sep2(mb_ctx *ctx, const mb_stencil_desc *first, const mb_stencil_desc *second,
                const int64_t *dims, const void *host_in, void *host_out);

/* iterateUntil (\n _ _ -> n >= n_steps-1) (\_ -> applyStencil ...)  (Manifest/Internal.hs:331-397):
 * the stencil is applied n_steps (>= 1) times, double-buffered; output size must equal input size.
 * dev_a holds the input; *result receives dev_a or dev_b, whichever holds the last array; BOTH arrays are
 * overwritten.  The host -> host form overlaps its copies with the passes where the stencil allows (the 3x3x3
 * star under Fill 0: a pass runs inside the dependency cone of what has arrived). */
int mb_iterate_dev(mb_ctx *ctx, const mb_stencil_desc *desc, const int64_t *dims, void *dev_a,
                   void *dev_b, int64_t n_steps, void **result);
int mb_iterate(mb_ctx *ctx, const mb_stencil_desc *desc, const int64_t *dims, const void *host_in,
               void *host_out, int64_t n_steps);

/* iterateUntil with a convergence predicate (Manifest/Internal.hs:331-349).  The reference's predicate is a closure
 * over the step number and the last two arrays; the device path evaluates the two predicates its users write:
 *   MB_UNTIL_EQUAL         cur == prev element-wise (Eq e: NaN /= NaN, -0 == 0): a fixed point, e.g. a Life board of
 *                          still lifes
 *   MB_UNTIL_MAX_ABS_DIFF  maximum |cur - prev| < tol (Float / Double; a NaN difference never converges)
 * after every `check_every`-th application (1: after every step, exactly iterateUntil; k > 1 is the predicate
 * `\n p c -> (n + 1) `mod` k == 0 && pred p c`), and gives up after max_steps applications.  The reduction runs on
 * the device; only its verdict travels to the host.  *steps receives the number of applications made, *converged
 * whether the predicate held, *result (device form) the buffer that holds the last array. */
typedef enum mb_until { MB_UNTIL_EQUAL = 0, MB_UNTIL_MAX_ABS_DIFF = 1 } mb_until;
int mb_iterate_until_dev(mb_ctx *ctx, const mb_stencil_desc *desc, const int64_t *dims, void *dev_a, void *dev_b,
                         int64_t max_steps, int64_t check_every, int predicate, double tol, int64_t *steps,
                         int32_t *converged, void **result);
int mb_iterate_until(mb_ctx *ctx, const mb_stencil_desc *desc, const int64_t *dims, const void *host_in, void *host_out,
                     int64_t max_steps, int64_t check_every, int predicate, double tol, int64_t *steps, int32_t *converged);

/* ---- slab-sharded iteration across the GPUs of one box (one process per GPU) ---------------- */
/* The global array is split along dims[0] into contiguous slabs, rank r owning planes
 * [r*N/R, (r+1)*N/R).  Every round (one step, or two fused steps for star-shaped 3-D kernels under
 * Fill 0) exchanges the ghost planes with rank±1: copy-engine peer copies over NVLink into IPC-mapped
 * staging areas with counter hand-shakes in device memory (no SM involved), or — when peer memory cannot
 * be mapped, or under option "no_p2p" / MASSIV_B200_NO_P2P — ncclSend/ncclRecv on a second stream.  NCCL
 * is always the rendezvous for the IPC handles.  The process rendezvous itself (who is who, the unique id
 * broadcast) is the host's job — torch.distributed in this repo. */
#define MB_NCCL_ID_BYTES 128
#define MB_MAX_HALO 8
#define MB_GHOST_FILL (-1)   /* ghost plane is the Fill element (physical border)            */
#define MB_GHOST_REMOTE (-2) /* ghost plane is received from the adjacent rank               */
/* What one rank owns and where its ghost planes come from (host-only arithmetic). */
typedef struct mb_shard_plan {
  int64_t first_plane;               /* global index of the rank's first plane: rank*N/R        */
  int64_t local_planes;              /* (rank+1)*N/R - first_plane                              */
  int64_t halo_lo, halo_hi;          /* ghost planes the buffers carry below / above the slab   */
  int32_t recv_lower_from;           /* rank my lower ghost planes come from, or -1               */
  int32_t recv_upper_from;           /* rank my upper ghost planes come from, or -1               */
  int32_t send_first_to;             /* rank that needs my first halo_hi planes (its upper ghost), or -1 */
  int32_t send_last_to;              /* rank that needs my last halo_lo planes (its lower ghost), or -1  */
  int64_t ghost_lo_src[MB_MAX_HALO]; /* per ghost plane: >= 0 own local plane to copy (Edge /   */
  int64_t ghost_hi_src[MB_MAX_HALO]; /* Reflect / Continue / 1-rank Wrap), MB_GHOST_FILL, or MB_GHOST_REMOTE */
  int64_t step_halo_lo, step_halo_hi; /* planes ONE step reads beyond the slab (pad_lo[0], pad_hi[0]: for
                                         mapStencil these are centre and size-1-centre);
                                         halo_* is twice that for stencils iterated two steps per
                                         pass (temporal blocking) */
} mb_shard_plan;
int mb_shard_plan_make(const mb_stencil_desc *desc, const int64_t *global_dims, int n_ranks, int rank,
                       mb_shard_plan *out);
int mb_nccl_unique_id(uint8_t id[MB_NCCL_ID_BYTES]);
int mb_shard_init(mb_ctx *ctx, int n_ranks, int rank, const uint8_t id[MB_NCCL_ID_BYTES]);
int mb_shard_halo(const mb_stencil_desc *desc, int64_t *halo_lo, int64_t *halo_hi);
/* local buffers dev_a/dev_b hold (halo_lo + local_dims[0] + halo_hi) planes: the rank's own
 * planes at plane offset halo_lo (ghost contents on entry are ignored).  global_dims[0] is the
 * global outer extent (for the border at the two ends of the global array).  *result receives
 * the BASE of the buffer (dev_a or dev_b) that holds the last state; its own planes again start
 * at plane offset halo_lo. */
int mb_iterate_sharded_dev(mb_ctx *ctx, const mb_stencil_desc *desc, const int64_t *global_dims,
                           const int64_t *local_dims, void *dev_a, void *dev_b, int64_t n_steps,
                           void **result);

/* how the last sharded call moved its halo planes: "p2p" (copy engines into IPC-mapped neighbour
 * memory, counters in device memory — no SM involved), "nccl" (ncclSend/ncclRecv) or "none" */
const char *mb_shard_transport(mb_ctx *ctx);
/* host -> host form: host_in / host_out hold this rank's OWN planes only (local_dims) */
int mb_iterate_sharded(mb_ctx *ctx, const mb_stencil_desc *desc, const int64_t *global_dims,
                       const int64_t *local_dims, const void *host_in, void *host_out, int64_t n_steps);

/* ---- device self-test (tests only) ----------------------------------------------------------- */
/* bitwise comparison of two device buffers: *mismatches = number of differing 8-byte words (bytes in an
 * unaligned tail), *first_offset = byte offset of the first one (UINT64_MAX when equal).  Used by the
 * full-size verification in bench.py and by the tests; blocks until the answer is known. */
int mb_compare_dev(mb_ctx *ctx, const void *dev_a, const void *dev_b, size_t bytes, uint64_t *mismatches,
                   uint64_t *first_offset);
/* compares the three-operation exact quotient avgStencil uses with the IEEE divide on `count`
 * random / adversarial bit patterns of element type MB_F32 or MB_F64; *mismatches must be 0 */
int mb_selftest_div(mb_ctx *ctx, int elem, int64_t n, uint64_t count, uint64_t seed,
                    uint64_t *mismatches, uint64_t *first_bad_bits);

/* ---- timing helpers (CUDA events on the context stream; used by bench.py) ------------------- */
int mb_timer_start(mb_ctx *ctx);
int mb_timer_stop(mb_ctx *ctx, float *milliseconds);

#ifdef __cplusplus
}
#endif
#endif /* MASSIV_B200_H */
