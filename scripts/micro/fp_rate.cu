// Micro-benchmark: issue rate of scalar vs packed (f32x2) unfused FP32 mul/add and of FP64 add/mul on
// sm_100a.  Informs how many unfused operations per cell a stencil can afford before it leaves the
// HBM roofline.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp_rate fp_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float *out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  float2 p0 = make_float2(x0, x1), p1 = make_float2(x2, x3), p2 = make_float2(x4, x5), p3 = make_float2(x6, x7);
  double d0 = x0, d1 = x1, d2 = x2, d3 = x3;
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {  // 8 independent chains of mul then add: 16 scalar ops
      x0 = __fadd_rn(__fmul_rn(x0, a), b); x1 = __fadd_rn(__fmul_rn(x1, a), b);
      x2 = __fadd_rn(__fmul_rn(x2, a), b); x3 = __fadd_rn(__fmul_rn(x3, a), b);
      x4 = __fadd_rn(__fmul_rn(x4, a), b); x5 = __fadd_rn(__fmul_rn(x5, a), b);
      x6 = __fadd_rn(__fmul_rn(x6, a), b); x7 = __fadd_rn(__fmul_rn(x7, a), b);
    } else if (MODE == 1) {  // same 16 flops as 8 packed ops
      p0 = __fadd2_rn(__fmul2_rn(p0, a2), b2); p1 = __fadd2_rn(__fmul2_rn(p1, a2), b2);
      p2 = __fadd2_rn(__fmul2_rn(p2, a2), b2); p3 = __fadd2_rn(__fmul2_rn(p3, a2), b2);
    } else if (MODE == 2) {  // 16 scalar adds
      x0 = __fadd_rn(x0, a); x1 = __fadd_rn(x1, a); x2 = __fadd_rn(x2, a); x3 = __fadd_rn(x3, a);
      x4 = __fadd_rn(x4, a); x5 = __fadd_rn(x5, a); x6 = __fadd_rn(x6, a); x7 = __fadd_rn(x7, a);
      x0 = __fadd_rn(x0, b); x1 = __fadd_rn(x1, b); x2 = __fadd_rn(x2, b); x3 = __fadd_rn(x3, b);
      x4 = __fadd_rn(x4, b); x5 = __fadd_rn(x5, b); x6 = __fadd_rn(x6, b); x7 = __fadd_rn(x7, b);
    } else if (MODE == 3) {  // 16 fused multiply-adds (for reference)
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    } else {  // 8 FP64 ops: 4 chains of mul then add
      d0 = __dadd_rn(__dmul_rn(d0, (double)a), (double)b); d1 = __dadd_rn(__dmul_rn(d1, (double)a), (double)b);
      d2 = __dadd_rn(__dmul_rn(d2, (double)a), (double)b); d3 = __dadd_rn(__dmul_rn(d3, (double)a), (double)b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + p0.x + p0.y + p1.x + p1.y + p2.x + p2.y +
                                               p3.x + p3.y + (float)(d0 + d1 + d2 + d3);
}

template <int MODE> void run(const char *name, int ops_per_iter) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float *out; cudaMalloc(&out, sms * 8 * 256 * sizeof(float));
  const int iters = 20000;
  k<MODE><<<sms * 8, 256>>>(out, 100, 0.999f, 0.001f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<sms * 8, 256>>>(out, iters, 0.999f, 0.001f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)sms * 8 * 256 * iters * ops_per_iter;
  printf("%-28s %8.2f Tops/s  (%.1f ops/clk/SM at 1.965 GHz)\n", name, ops / ms / 1e9, ops / ms / 1e9 * 1e3 / 1.965 / sms * 1e-3 * 1e3 / 1e3);
  cudaFree(out);
}

int main() {
  run<0>("scalar mul+add (unfused)", 16);
  run<1>("packed f32x2 mul+add", 16);
  run<2>("scalar add", 16);
  run<3>("scalar fma (1 op each)", 16);
  run<4>("fp64 mul+add (unfused)", 8);
  return 0;
}
