/*
 * massiv_oracle.h — CPU ORACLE for the massiv stencil / windowed-load path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (massiv_b200/, include/massiv_b200.h) never links, imports or calls anything here.
 *
 * It is a plain-C restatement of the reference algorithm (lehins/massiv 1.0.5.0, all
 * Haskell; no GHC exists in the build image so the reference itself cannot be run).
 * Parity status: PINNED against the reference's own golden vectors (tests/test_oracle_golden.py
 * replays massiv-test/tests/Test/Massiv/Array/StencilSpec.hs and the Stencil.hs doctests).
 *
 * Reference lines restated (paths under /root/reference/massiv/src/Data/Massiv/):
 *   geometry      Array/Stencil.hs:173-179 (samePadding), :197-217 (applyStencil)
 *   borders       Core/Index.hs:236-253 (handleBorderIndex),
 *                 Core/Index/Internal.hs:653-658, 1030-1035 (repairIndex, applied once per dim)
 *   fold stencils Array/Stencil.hs:298-302, 335-337, 368-370, 401-403, 425-427, 454-456, 479-481
 *   conv / corr   Array/Stencil/Convolution.hs:44-57, 64-80, 89-100, 107-121
 *   algebra       Array/Stencil/Internal.hs:83-112 (size/centre union), :127-146
 *   stride        Core/Index/Stride.hs:91-120, Array/Delayed/Windowed.hs:245-280, 307-337, 374-430
 *   iterate       Array/Manifest/Internal.hs:331-397
 *   Life rule     massiv-examples/GameOfLife/app/GameOfLife.hs:15-36
 */
#ifndef MASSIV_ORACLE_H
#define MASSIV_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MO_MAX_RANK 5

enum { MO_OK = 0, MO_ERR_INVALID_DESC = 1, MO_ERR_SIZE_MISMATCH = 2, MO_ERR_INDEX_ZERO = 3,
       MO_ERR_UNSUPPORTED = 4 };

/* element types: sizes follow Storable.sizeOf (Manifest/Storable.hs:71-77): Bool is 4 bytes */
enum { MO_U8 = 0, MO_I8, MO_U16, MO_I16, MO_U32, MO_I32, MO_U64, MO_I64, MO_F32, MO_F64, MO_BOOL32 };

enum { MO_ID = 0,      /* idStencil                                  Stencil.hs:269          */
       MO_SUM,         /* sumStencil      acc0 = 0, acc = acc + x    Stencil.hs:425-427      */
       MO_PRODUCT,     /* productStencil  acc0 = 1, acc = acc * x    Stencil.hs:454-456      */
       MO_AVG,         /* avgStencil      sum / fromIntegral n       Stencil.hs:479-481      */
       MO_MAX,         /* maxStencil      acc0 = minBound            Stencil.hs:368-370      */
       MO_MIN,         /* minStencil      acc0 = maxBound            Stencil.hs:401-403      */
       MO_CONV,        /* makeConvolutionStencilFromKernel           Convolution.hs:64-80    */
       MO_CORR,        /* makeCorrelationStencilFromKernel           Convolution.hs:107-121  */
       MO_TAPS,        /* make(Unsafe){Convolution,Correlation}Stencil closure form: an ordered
                          tap list, acc = src[ix + off_t] * k_t + acc  Convolution.hs:44-57,89-100 */
       MO_LIFE,        /* Game of Life rule stencil                  GameOfLife.hs:15-36     */
       MO_MAG2,        /* sqrt (fmap (^2) a + fmap (^2) b), a and b CONV (or CORR with flag)
                          kernels of identical size                  Bench/Sobel.hs:39-44    */
       /* Integral approximation rules: 1-D along one dimension, centre 0, coeffs[0] = dx  */
       MO_INT_MIDPOINT,  /* midpointStencil                          Numeric/Integral.hs:54-66   */
       MO_INT_TRAPEZOID, /* trapezoidStencil                         Numeric/Integral.hs:75-90   */
       MO_INT_SIMPSON,   /* simpsonsStencil (n even)                 Numeric/Integral.hs:99-118  */
       MO_EXPR           /* Applicative / Num / Fractional / Floating instances of Stencil: terms (sub-stencils of
                            kinds MO_ID..MO_TAPS) combined point-wise by a postfix program; size / centre are the
                            unions of Stencil/Internal.hs:83-90                     Stencil/Internal.hs:93-174  */
};

/* MO_EXPR: a term is a sub-stencil evaluated at the SAME absolute source index as every other term
 * (stF ug gV ix = f (f1 ug gV ix) (f2 ug gV ix), Internal.hs:104-111) */
typedef struct mo_term {
  int32_t kind, reserved;
  int64_t size[MO_MAX_RANK], center[MO_MAX_RANK];
  const void *coeffs;
  int64_t ntaps;
  const int64_t *tap_offsets;
} mo_term;
enum { MO_X_TERM = 1, MO_X_CONST, MO_X_ADD, MO_X_SUB, MO_X_MUL, MO_X_DIV, MO_X_MIN, MO_X_MAX, MO_X_NEGATE, MO_X_ABS,
       MO_X_SIGNUM, MO_X_SQUARE, MO_X_RECIP, MO_X_SQRT };
typedef struct mo_xins { int32_t op, arg; uint8_t operand[8]; } mo_xins;
#define MO_MAX_TERMS 8
#define MO_MAX_PROGRAM 32
#define MO_MAX_STACK 8

enum { MO_FILL = 0, MO_WRAP, MO_EDGE, MO_REFLECT, MO_CONTINUE };

#define MO_FLAG_MAG2_CORR 1u   /* MAG2 sub-stencils are correlations instead of convolutions */
#define MO_FLAG_NO_FAST (1u << 31) /* oracle only: disable the vectorised interior loop nest */

/* `fmap f` over the stencil's results: rmapStencil (Stencil/Internal.hs:70-83), the Num / Fractional /
 * Floating methods defined through it (:114-146), Functor (Array DW ix) (Delayed/Windowed.hs:78-84).
 * Only f with an exactly specified result; semantics are GHC's Prelude instances (GHC.Float, GHC.Num,
 * GHC.Classes default min/max).  No reference test pins these beyond the Prelude's own definitions. */
enum { MO_POST_NEGATE = 1, MO_POST_ABS, MO_POST_SIGNUM, MO_POST_SQUARE, MO_POST_ADD, MO_POST_SUB, MO_POST_RSUB,
       MO_POST_MUL, MO_POST_MIN, MO_POST_MAX, MO_POST_DIV, MO_POST_RDIV, MO_POST_RECIP, MO_POST_SQRT };
#define MO_MAX_POST 8
typedef struct mo_post_op { int32_t op; int32_t reserved; uint8_t operand[8]; } mo_post_op;

/* Field-for-field the same layout as mb_stencil_desc in include/massiv_b200.h so the tests can
 * hand one ctypes structure to both sides; the two headers are otherwise independent. */
typedef struct mo_desc {
  int32_t kind, elem, rank, border;
  int64_t size[MO_MAX_RANK];     /* stencil size, outermost dimension first              */
  int64_t center[MO_MAX_RANK];   /* stencilCenter (already "inverted" for convolutions)  */
  int64_t pad_lo[MO_MAX_RANK];   /* paddingFromOrigin                                    */
  int64_t pad_hi[MO_MAX_RANK];   /* paddingFromBottom                                    */
  int64_t stride[MO_MAX_RANK];   /* computeWithStride; 1 everywhere for compute          */
  uint8_t fill[8];               /* Fill value, native-endian element                    */
  const void *coeffs;            /* CONV/CORR/MAG2: prod(size) elements row-major; TAPS: ntaps; INT_*: dx */
  const void *coeffs2;           /* MAG2 second kernel                                   */
  int64_t ntaps;                 /* TAPS only                                            */
  const int64_t *tap_offsets;    /* TAPS only: ntaps*rank source offsets, application order */
  uint32_t flags;
  uint32_t n_post;               /* post-maps (fmap f over the results), applied in order        */
  const struct mo_post_op *post;
  uint32_t n_terms, n_program;   /* MO_EXPR                                                       */
  const mo_term *terms;
  const mo_xins *program;
} mo_desc;

/* repaired index of handleBorderIndex for one dimension; *fill set to 1 when a Fill border
 * yields the fill element instead.  Returns MO_ERR_INDEX_ZERO for k <= 0 and non-Fill. */
int mo_border_index(int32_t border, int64_t k, int64_t i, int64_t *out, int32_t *fill);

/* stencil geometry (size used for the output-size arithmetic) and expected centre */
int mo_geom(const mo_desc *d, int64_t *geom_size, int64_t *expect_center);

/* output dims of compute(WithStride) (applyStencil padding stencil arr) */
int mo_out_dims(const mo_desc *d, const int64_t *in_dims, int64_t *out_dims);

/* compute(WithStride) . applyStencil;  nthreads mimics Par (rows / outer pages split) */
int mo_apply(const mo_desc *d, const int64_t *in_dims, const void *in, void *out, int nthreads);

/* iterateUntil (\n _ _ -> n >= n_steps-1) (\_ -> applyStencil ...) : requires out dims == in dims.
 * tmp is a scratch buffer of the same size; result always lands in `out`. */
int mo_iterate(const mo_desc *d, const int64_t *dims, const void *in, void *out, void *tmp,
               int64_t n_steps, int nthreads);

int mo_elem_size(int32_t elem);

#ifdef __cplusplus
}
#endif
#endif
