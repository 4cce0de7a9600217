/*
 * Plain-C driver of libmassiv_b200.so — the calls the Haskell shim makes, with no Python, torch or
 * C++ in the process.  Built and run by tests/test_cabi_harness.py:
 *     cabi_harness geometry   host-only entry points (runs anywhere)
 *     cabi_harness gpu        golden vectors of the reference through mb_run / mb_iterate (needs a GPU)
 * Expected values are literal transcriptions of massiv's own tests (see tests/golden/).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "massiv_b200.h"

static int failures = 0;
#define CHECK(cond, msg)                                        \
  do {                                                          \
    if (!(cond)) { printf("FAIL %s (%s:%d)\n", msg, __FILE__, __LINE__); ++failures; } \
  } while (0)

static void desc_init(mb_stencil_desc *d, int kind, int elem, int rank, int border) {
  memset(d, 0, sizeof(*d));
  d->kind = kind; d->elem = elem; d->rank = rank; d->border = border;
  for (int i = 0; i < rank; ++i) d->stride[i] = 1;
}

/* mapStencil: pad_lo = centre, pad_hi = size - (centre + 1) */
static void same_padding(mb_stencil_desc *d) {
  for (int i = 0; i < d->rank; ++i) {
    d->pad_lo[i] = d->center[i];
    d->pad_hi[i] = d->size[i] - (d->center[i] + 1);
    if (d->pad_hi[i] < 0) d->pad_hi[i] = 0;
  }
}

static int geometry(void) {
  mb_stencil_desc d;
  int64_t in[2] = {2, 3}, out[2];
  /* idStencil with Padding (1,2) (3,4): Stencil.hs:102-110 -> 6 x 9 */
  desc_init(&d, MB_ID, MB_I64, 2, MB_FILL);
  d.size[0] = d.size[1] = 1;
  d.pad_lo[0] = 1; d.pad_lo[1] = 2; d.pad_hi[0] = 3; d.pad_hi[1] = 4;
  CHECK(mb_out_dims(&d, in, out) == MB_OK && out[0] == 6 && out[1] == 9, "idStencil padded size");
  /* maxStencil (Sz 3) with Stride 3 on 9x9 -> 3x3: Stencil.hs:360-365 */
  desc_init(&d, MB_MAX, MB_I64, 2, MB_EDGE);
  d.size[0] = d.size[1] = 3; same_padding(&d); d.stride[0] = d.stride[1] = 3;
  in[0] = in[1] = 9;
  CHECK(mb_out_dims(&d, in, out) == MB_OK && out[0] == 3 && out[1] == 3, "strided size");
  /* Bounded has no Double instance */
  d.elem = MB_F64;
  CHECK(mb_out_dims(&d, in, out) == MB_ERR_UNSUPPORTED, "maxStencil on Double is rejected");
  /* handleBorderIndex doctests: Core/Index.hs:218-223 */
  int64_t r; int32_t f;
  CHECK(mb_border_index(MB_WRAP, 3, 3, &r, &f) == MB_OK && r == 0 && !f, "Wrap");
  CHECK(mb_border_index(MB_EDGE, 2, 2, &r, &f) == MB_OK && r == 1, "Edge");
  CHECK(mb_border_index(MB_FILL, 2, 2, &r, &f) == MB_OK && f == 1, "Fill");
  CHECK(mb_border_index(MB_REFLECT, 0, 1, &r, &f) == MB_ERR_INDEX_ZERO, "IndexZeroException");
  CHECK(mb_elem_size(MB_BOOL32) == 4 && mb_version() == MB_VERSION, "sizes / version");
  return failures;
}

static int gpu(void) {
  mb_ctx *ctx = NULL;
  int st = mb_ctx_create(0, &ctx);
  if (st != MB_OK) { printf("FAIL mb_ctx_create -> %d (%s): no CPU fallback exists\n", st, mb_status_string(st)); return 1; }
  mb_stencil_desc d;
  /* StencilSpec.hs:227: conv [1,2,3] over [1..5], Fill 0 -> [4,10,16,22,22] */
  {
    int64_t ys[5] = {1, 2, 3, 4, 5}, k[3] = {1, 2, 3}, out[5], want[5] = {4, 10, 16, 22, 22}, dims[1] = {5};
    desc_init(&d, MB_CONV, MB_I64, 1, MB_FILL);
    d.size[0] = 3; d.center[0] = 1; d.coeffs = k; same_padding(&d);
    st = mb_run(ctx, &d, dims, ys, out);
    CHECK(st == MB_OK && memcmp(out, want, sizeof(want)) == 0, "convolution known answer");
    /* :237 correlation -> [8,14,20,26,14] */
    int64_t wantc[5] = {8, 14, 20, 26, 14};
    d.kind = MB_CORR;
    st = mb_run(ctx, &d, dims, ys, out);
    CHECK(st == MB_OK && memcmp(out, wantc, sizeof(wantc)) == 0, "correlation known answer");
  }
  /* Stencil.hs:470-474: avgStencil (Sz2 2 3) noPadding on 11..22 -> [[14,15],[18,19]] */
  {
    double a[12], out[4], want[4] = {14, 15, 18, 19};
    int64_t dims[2] = {3, 4};
    for (int i = 0; i < 12; ++i) a[i] = 11 + i;
    desc_init(&d, MB_AVG, MB_F64, 2, MB_EDGE);
    d.size[0] = 2; d.size[1] = 3;
    st = mb_run(ctx, &d, dims, a, out);
    CHECK(st == MB_OK && memcmp(out, want, sizeof(want)) == 0, "avgStencil doctest");
  }
  /* Life: a blinker has period 2 on a torus (GameOfLife.hs:15-36, 47-52) */
  {
    uint8_t g[64 * 64], out[64 * 64];
    int64_t dims[2] = {64, 64};
    memset(g, 0, sizeof(g));
    g[10 * 64 + 20] = g[11 * 64 + 20] = g[12 * 64 + 20] = 1;
    desc_init(&d, MB_LIFE, MB_U8, 2, MB_WRAP);
    d.size[0] = d.size[1] = 3; d.center[0] = d.center[1] = 1; same_padding(&d);
    st = mb_iterate(ctx, &d, dims, g, out, 1);
    CHECK(st == MB_OK && out[11 * 64 + 19] == 1 && out[11 * 64 + 20] == 1 && out[11 * 64 + 21] == 1 && out[10 * 64 + 20] == 0,
          "blinker after one generation");
    st = mb_iterate(ctx, &d, dims, g, out, 100);
    CHECK(st == MB_OK && memcmp(out, g, sizeof(g)) == 0, "blinker after 100 generations");
    d.pad_hi[0] = 0; /* no longer size-preserving: iterate must refuse */
    st = mb_iterate(ctx, &d, dims, g, out, 2);
    CHECK(st == MB_ERR_SIZE_MISMATCH, "iterate refuses a shrinking stencil");
  }
  /* IndexZeroException through the compute entry point */
  {
    int64_t dims[2] = {0, 3}, out[3];
    desc_init(&d, MB_SUM, MB_I64, 2, MB_EDGE);
    d.size[0] = d.size[1] = 2; d.pad_lo[0] = 1; d.pad_hi[0] = 1; d.pad_hi[1] = 1;
    st = mb_run(ctx, &d, dims, out, out);
    CHECK(st == MB_ERR_INDEX_ZERO && strlen(mb_last_error(ctx)) > 0, "IndexZeroException status");
  }
  printf("launches=%lld last_kernel=%s\n", (long long)mb_launch_count(ctx), mb_last_kernel(ctx));
  CHECK(mb_launch_count(ctx) > 0, "kernels were launched");
  mb_ctx_destroy(ctx);
  return failures;
}

int main(int argc, char **argv) {
  const char *mode = argc > 1 ? argv[1] : "geometry";
  int f = strcmp(mode, "gpu") == 0 ? gpu() : geometry();
  printf(f ? "CABI_HARNESS_FAIL %d\n" : "CABI_HARNESS_OK\n", f);
  return f ? 1 : 0;
}
