// Post-maps: `fmap f` over a stencil's results (Functor (Stencil ix e) / Functor (Array DW ix),
// Stencil/Internal.hs:36-41, Delayed/Windowed.hs:78-84) for the f listed in mb_post — each one an
// exactly specified IEEE-754 / two's-complement operation, so the mapped result stays bit-identical.
// First version: an in-place elementwise pass over what the stencil kernel wrote (one extra read and
// write of the OUTPUT per launch); fusing it into the kernels' store epilogues is the follow-up.
#include "mb_post.cuh"

namespace mb {

template <typename T>
__global__ void __launch_bounds__(256) post_kernel(const PostParams P, T *__restrict__ data, size_t n) {
  T c[MB_MAX_POST];
#pragma unroll
  for (int i = 0; i < MB_MAX_POST; ++i) {
    unsigned long long b = P.operand[i];
    memcpy(&c[i], &b, sizeof(T));
  }
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    T x = data[i];
#pragma unroll
    for (int k = 0; k < MB_MAX_POST; ++k)
      if (k < P.n) x = post_one<T>(P.op[k], x, c[k]);
    data[i] = x;
  }
}

template <typename T>
static int launch_post_t(Ctx *c, const PostParams &P, void *data, size_t n) {
  size_t blocks = (n + 255) / 256;
  const size_t cap = (size_t)c->num_sms * 16;
  if (blocks > cap) blocks = cap;
  post_kernel<T><<<(unsigned)blocks, 256, 0, c->stream>>>(P, (T *)data, n);
  c->launches++;
  c->aux_launches++;
  return check_cuda(c, cudaGetLastError(), "post_kernel launch");
}

int launch_post(Ctx *c, const Plan &p, void *d_out, const OuterRange *range) {
  if (p.post.empty()) return MB_OK;
  const PostParams P = make_post_params(p);
  int64_t plane = 1;
  for (int i = 1; i < p.rank; ++i) plane *= p.out_dims[i];
  int64_t o0 = 0, o1 = p.out_dims[0];
  if (range) { o0 = range->o0; o1 = range->o1; }
  if (o1 <= o0 || plane == 0) return MB_OK;
  char *base = (char *)d_out + (size_t)o0 * plane * p.esize;
  const size_t n = (size_t)(o1 - o0) * plane;
  switch (p.elem) {
    case MB_U8: return launch_post_t<uint8_t>(c, P, base, n);
    case MB_I8: return launch_post_t<int8_t>(c, P, base, n);
    case MB_U16: return launch_post_t<uint16_t>(c, P, base, n);
    case MB_I16: return launch_post_t<int16_t>(c, P, base, n);
    case MB_U32: return launch_post_t<uint32_t>(c, P, base, n);
    case MB_I32: return launch_post_t<int32_t>(c, P, base, n);
    case MB_U64: return launch_post_t<uint64_t>(c, P, base, n);
    case MB_I64: return launch_post_t<int64_t>(c, P, base, n);
    case MB_F32: return launch_post_t<float>(c, P, base, n);
    case MB_F64: return launch_post_t<double>(c, P, base, n);
  }
  return fail(c, MB_ERR_UNSUPPORTED, "post-maps on this element type");
}

}  // namespace mb
