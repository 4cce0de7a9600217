// Game of Life rule stencil on Word8 cells (massiv-examples/GameOfLife/app/GameOfLife.hs:15-36):
//     n = sum of the 8 neighbours (Word8, modulo 256);  out = 1 if (c == 0 and n == 3) or
//     (c == 1 and (n == 2 or n == 3)) else 0  — any centre value other than 0/1 yields 0.
//
// Two kernels:
//   life_byte_kernel  one generation, byte-exact for ARBITRARY cell values (mod-256 sums), every
//                     border.  Four cells per 32-bit word, SWAR arithmetic, rows reused vertically.
//   life_bits_kernel  many generations in ONE launch.  After one byte-exact generation every cell is
//                     0 or 1, so the grid is packed to one bit per cell (2048^2 -> 512 KiB, resident
//                     in shared memory across the SMs) and advanced with bit-sliced adders, 64 cells
//                     per operation.  Temporal blocking: each CTA owns a strip of rows plus T halo
//                     rows, runs T generations out of shared memory, then trades boundary rows with
//                     its two neighbours through L2 with flag hand-shakes (no grid-wide barrier, no
//                     relaunch).  The per-generation cost drops from a kernel launch (~2-3 us, far
//                     above the 1.3 us HBM time of config 1) to well under 1 us.
// Both are bit-exact against the reference rule for every input (tests/test_gpu_parity.py).
#include <cooperative_groups.h>

#include "mb_arith.cuh"
#include "mb_internal.h"

namespace mb {

// ---- byte-wise generation ------------------------------------------------------------------------

struct LifeByteParams {
  int h, w, border;
  unsigned fill;
};

// per-byte wrapping add of four packed cells
__device__ __forceinline__ unsigned add4(unsigned a, unsigned b) {
  return ((a & 0x7f7f7f7fu) + (b & 0x7f7f7f7fu)) ^ ((a ^ b) & 0x80808080u);
}
// 0x01 in every byte of x that equals the byte k, else 0x00 (exact, no cross-byte borrow)
__device__ __forceinline__ unsigned eq4(unsigned x, unsigned k) {
  const unsigned y = x ^ (k * 0x01010101u);
  return (~(((y & 0x7f7f7f7fu) + 0x7f7f7f7fu) | y) & 0x80808080u) >> 7;
}

// row y of the source, as the centre word and its left/right neighbour words, border-resolved
__device__ __forceinline__ void life_row(const LifeByteParams &P, const unsigned char *__restrict__ in, int y, int x0,
                                         unsigned &l, unsigned &c, unsigned &r) {
  const unsigned f4 = P.fill * 0x01010101u;
  int yy = y;
  if (!resolve_border32(P.border, P.h, yy)) { l = c = r = f4; return; }
  const unsigned char *row = in + (size_t)yy * P.w;
  c = *reinterpret_cast<const unsigned *>(row + x0);
  int xl = x0 - 1, xr = x0 + 4;
  const unsigned lb = resolve_border32(P.border, P.w, xl) ? row[xl] : P.fill;
  const unsigned rb = resolve_border32(P.border, P.w, xr) ? row[xr] : P.fill;
  l = (c << 8) | lb;          // cell x-1 for each of the four cells (byte 0 is the lowest x)
  r = (c >> 8) | (rb << 24);  // cell x+1
}

constexpr int LIFE_R = 4;  // rows per thread

__global__ void __launch_bounds__(256) life_byte_kernel(const LifeByteParams P, const unsigned char *__restrict__ in,
                                                        unsigned char *__restrict__ out) {
  const int wx = P.w >> 2;
  const int gx = blockIdx.x * blockDim.x + threadIdx.x;
  if (gx >= wx) return;
  const int x0 = gx << 2;
  const int y0 = blockIdx.y * LIFE_R;
  unsigned l, c, r;
  life_row(P, in, y0 - 1, x0, l, c, r);
  unsigned t_up = add4(add4(l, c), r);
  life_row(P, in, y0, x0, l, c, r);
  unsigned t_mid = add4(add4(l, c), r), m_mid = add4(l, r), c_mid = c;
#pragma unroll
  for (int i = 0; i < LIFE_R; ++i) {
    const int y = y0 + i;
    if (y >= P.h) return;
    life_row(P, in, y + 1, x0, l, c, r);
    const unsigned t_dn = add4(add4(l, c), r);
    const unsigned n = add4(add4(t_up, t_dn), m_mid);
    const unsigned n3 = eq4(n, 3), n2 = eq4(n, 2), c0 = eq4(c_mid, 0), c1 = eq4(c_mid, 1);
    *reinterpret_cast<unsigned *>(out + (size_t)y * P.w + x0) = (c0 & n3) | (c1 & (n2 | n3));
    t_up = t_mid;
    t_mid = t_dn;
    m_mid = add4(l, r);
    c_mid = c;
  }
}

static bool life_byte_ok(const Plan &p, const void *a, const void *b) {
  return p.kind == MB_LIFE && p.rank == 2 && p.unit_stride() && p.same_shape() && p.in_dims[1] % 4 == 0 &&
         p.in_dims[0] >= 1 && p.in_dims[1] >= 4 && p.in_dims[0] < (1ll << 30) && p.in_dims[1] < (1ll << 30) &&
         ((uintptr_t)a % 4 == 0) && ((uintptr_t)b % 4 == 0);
}

int launch_life(Ctx *c, const DevPlan &dp, const void *d_in, void *d_out) {
  const Plan &p = dp.p;
  if (!life_byte_ok(p, d_in, d_out)) return -1;
  LifeByteParams P;
  P.h = (int)p.in_dims[0]; P.w = (int)p.in_dims[1]; P.border = p.border; P.fill = p.fill[0];
  const int wx = P.w / 4;
  dim3 block(wx < 256 ? ((wx + 31) / 32) * 32 : 256);
  dim3 grid((wx + block.x - 1) / block.x, (P.h + LIFE_R - 1) / LIFE_R);
  life_byte_kernel<<<grid, block, 0, c->stream>>>(P, (const unsigned char *)d_in, (unsigned char *)d_out);
  c->launches++;
  c->last_kernel = "life_byte";
  return check_cuda(c, cudaGetLastError(), "life_byte_kernel launch");
}

// ---- bit-packed, temporally blocked, many generations per launch ----------------------------------

struct LifeBitsParams {
  int h, w, wpr;        // wpr = 64-bit words per row
  int rows_own, T;      // rows per CTA strip, generations per block (= halo rows)
  int wrap;             // 1: Wrap, 0: Fill 0
  int ngen;             // generations to advance
  unsigned long long *bound;  // [cta][parity][side][T][wpr] published boundary rows
  int *flags;                 // [cta] number of completed blocks
};

// one generation of row `mid` given the rows above/below, for word column w (bit i = cell 64*w + i)
__device__ __forceinline__ unsigned long long life_word(unsigned long long ul, unsigned long long uc, unsigned long long ur,
                                                        unsigned long long ml, unsigned long long mc, unsigned long long mr,
                                                        unsigned long long dl, unsigned long long dc, unsigned long long dr) {
  // neighbours of bit i: bit i-1 / i+1 of the same word, carried across word boundaries
  const unsigned long long uL = (uc << 1) | (ul >> 63), uR = (uc >> 1) | (ur << 63);
  const unsigned long long mL = (mc << 1) | (ml >> 63), mR = (mc >> 1) | (mr << 63);
  const unsigned long long dL = (dc << 1) | (dl >> 63), dR = (dc >> 1) | (dr << 63);
  // bit-sliced adders: rows above and below are 3-input (2-bit result), the middle row 2-input
  const unsigned long long t0 = uL ^ uc ^ uR, t1 = (uL & uc) | (uR & (uL ^ uc));
  const unsigned long long b0 = dL ^ dc ^ dR, b1 = (dL & dc) | (dR & (dL ^ dc));
  const unsigned long long m0 = mL ^ mR, m1 = mL & mR;
  const unsigned long long s0 = t0 ^ b0 ^ m0, k0 = (t0 & b0) | (m0 & (t0 ^ b0));  // ones, carry into twos
  const unsigned long long x = t1 ^ b1 ^ m1, k1 = (t1 & b1) | (m1 & (t1 ^ b1));   // twos before carry, carry into fours
  const unsigned long long s1 = x ^ k0, k2 = x & k0;
  const unsigned long long fours = k1 ^ k2;  // (k1 & k2 means n == 8, where s1 is 0 anyway)
  // n == 3 -> alive; n == 2 -> unchanged; otherwise dead
  return s1 & ~fours & ~(k1 & k2) & (s0 | mc);
}

__global__ void __launch_bounds__(256, 1) life_bits_kernel(const LifeBitsParams P, const unsigned char *__restrict__ in,
                                                           unsigned char *__restrict__ out) {
  extern __shared__ __align__(16) unsigned long long lsm[];
  const int T = P.T, wpr = P.wpr, rows = P.rows_own + 2 * T;
  unsigned long long *cur = lsm, *nxt = lsm + (size_t)rows * wpr;
  const int tid = threadIdx.x, cta = blockIdx.x, ncta = gridDim.x;
  const int row0 = cta * P.rows_own - T;  // global row of local row 0
  const int up = (cta + ncta - 1) % ncta, down = (cta + 1) % ncta;
  const size_t bstride = (size_t)T * wpr;  // words per (cta, parity, side)

  // ---- pack: bytes (0/1 after the byte-exact first generation) -> bits, own strip + halos ----
  for (int e = tid; e < rows * wpr; e += 256) {
    const int lr = e / wpr, w = e - lr * wpr;
    int gy = row0 + lr;
    unsigned long long bits = 0;
    bool inside = true;
    if (P.wrap) gy = floor_mod_dev32(gy, P.h); else inside = gy >= 0 && gy < P.h;
    if (inside) {
      const uint4 *src = reinterpret_cast<const uint4 *>(in + (size_t)gy * P.w + (size_t)w * 64);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 v = __ldcg(src + q);
        const unsigned wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // gather bit 0 of each of the four bytes into a nibble
          const unsigned b = wd[k] & 0x01010101u;
          const unsigned nib = (b | (b >> 7) | (b >> 14) | (b >> 21)) & 0xFu;
          bits |= (unsigned long long)nib << (q * 16 + k * 4);
        }
      }
    }
    cur[e] = bits;
  }
  __syncthreads();

  const int nblocks = (P.ngen + T - 1) / T;
  // thread -> (word column, contiguous chunk of rows): the three-row window slides down in registers
  const int col = tid % wpr;
  const int lanes_y = 256 / wpr;  // wpr is a power of two <= 256 (checked on the host)
  const int ty = tid / wpr;
  const int chunk = (rows - 2 + lanes_y - 1) / lanes_y;
  const int rbeg = 1 + ty * chunk, rend = min(rows - 1, rbeg + chunk);
  const int cl = (col + wpr - 1) % wpr, cr = (col + 1) % wpr;
  const bool lwrap = P.wrap || col > 0, rwrap = P.wrap || col < wpr - 1;

  for (int blk = 0; blk < nblocks; ++blk) {
    if (blk > 0) {
      // wait for both neighbours to have published their state after block blk-1, then take halos
      if (tid == 0) {
        volatile int *f = P.flags;
        while (f[up] < blk || f[down] < blk) __nanosleep(64);
        __threadfence();
      }
      __syncthreads();
      const int par = blk & 1;
      const unsigned long long *from_up = P.bound + (((size_t)up * 2 + par) * 2 + 1) * bstride;    // their bottom rows
      const unsigned long long *from_dn = P.bound + (((size_t)down * 2 + par) * 2 + 0) * bstride;  // their top rows
      const bool has_up = P.wrap || cta > 0, has_dn = P.wrap || cta < ncta - 1;
      for (int e = tid; e < T * wpr; e += 256) {
        cur[e] = has_up ? __ldcg(from_up + e) : 0ull;
        cur[(size_t)(T + P.rows_own) * wpr + e] = has_dn ? __ldcg(from_dn + e) : 0ull;
      }
      __syncthreads();
    }
    const int gens = min(T, P.ngen - blk * T);
    for (int g = 0; g < gens; ++g) {
      if (rbeg < rend) {
        unsigned long long ul = lwrap ? cur[(rbeg - 1) * wpr + cl] : 0, uc = cur[(rbeg - 1) * wpr + col],
                           ur = rwrap ? cur[(rbeg - 1) * wpr + cr] : 0;
        unsigned long long ml = lwrap ? cur[rbeg * wpr + cl] : 0, mc = cur[rbeg * wpr + col],
                           mr = rwrap ? cur[rbeg * wpr + cr] : 0;
        for (int r = rbeg; r < rend; ++r) {
          const unsigned long long dl = lwrap ? cur[(r + 1) * wpr + cl] : 0, dc = cur[(r + 1) * wpr + col],
                                   dr = rwrap ? cur[(r + 1) * wpr + cr] : 0;
          unsigned long long o = life_word(ul, uc, ur, ml, mc, mr, dl, dc, dr);
          if (!P.wrap) {  // Fill 0: rows outside the grid stay dead
            const int gy = row0 + r;
            if (gy < 0 || gy >= P.h) o = 0;
          }
          nxt[r * wpr + col] = o;
          ul = ml; uc = mc; ur = mr;
          ml = dl; mc = dc; mr = dr;
        }
      }
      if (tid < wpr) {  // outermost halo rows carry no information any more; keep them defined
        nxt[tid] = 0;
        nxt[(size_t)(rows - 1) * wpr + tid] = 0;
      }
      __syncthreads();
      unsigned long long *t = cur; cur = nxt; nxt = t;
    }
    if (blk + 1 < nblocks) {
      // publish my top and bottom T rows (state after this block) for the neighbours' next block
      const int par = (blk + 1) & 1;
      unsigned long long *top = P.bound + (((size_t)cta * 2 + par) * 2 + 0) * bstride;
      unsigned long long *bot = P.bound + (((size_t)cta * 2 + par) * 2 + 1) * bstride;
      for (int e = tid; e < T * wpr; e += 256) {
        top[e] = cur[(size_t)T * wpr + e];
        bot[e] = cur[(size_t)P.rows_own * wpr + e];
      }
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        *((volatile int *)&P.flags[cta]) = blk + 1;
        __threadfence();
      }
    }
  }

  // ---- unpack the own strip: bits -> Word8 cells ----
  for (int e = tid; e < P.rows_own * wpr; e += 256) {
    const int lr = e / wpr, w = e - lr * wpr;
    const unsigned long long bits = cur[(size_t)(T + lr) * wpr + w];
    uint4 *dst = reinterpret_cast<uint4 *>(out + (size_t)(cta * P.rows_own + lr) * P.w + (size_t)w * 64);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      unsigned wd[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned nib = (unsigned)(bits >> (q * 16 + k * 4)) & 0xFu;
        wd[k] = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
      }
      dst[q] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
    }
  }
}

int life_iterate(Ctx *c, const DevPlan &dp, void *d_a, void *d_b, int64_t n_steps, void **result) {
  const Plan &p = dp.p;
  if (!c->life_bitpack || !life_byte_ok(p, d_a, d_b) || n_steps < 3) return -1;
  if (n_steps - 1 > (int64_t)0x7fffffff) return -1;  // the kernel counts generations in an int: single steps instead
  const int h = (int)p.in_dims[0], w = (int)p.in_dims[1];
  bool fill0 = p.border == MB_FILL && p.fill[0] == 0;
  if (p.border != MB_WRAP && !fill0) return -1;
  if (w % 64 != 0 || ((uintptr_t)d_a % 16) || ((uintptr_t)d_b % 16)) return -1;
  const int wpr = w / 64;
  if (wpr > 256 || (wpr & (wpr - 1)) != 0) return -1;
  int rows_own = 0;
  for (int cand : {16, 32, 8, 64}) {
    if (h % cand == 0 && h / cand <= c->num_sms) { rows_own = cand; break; }
  }
  if (!rows_own) return -1;
  const int T = rows_own < 16 ? rows_own : 16;
  const int ncta = h / rows_own;
  const size_t smem = (size_t)2 * (rows_own + 2 * T) * wpr * sizeof(unsigned long long);
  if (smem > 200 * 1024) return -1;
  // generation 1 byte-exact (cells may hold any Word8), a -> b; the rest bit-packed, b -> a
  int st = launch_life(c, dp, d_a, d_b);
  if (st) return st;
  const size_t need = (size_t)ncta * 2 * 2 * T * wpr;
  if (need > c->life_bound_words) {
    MB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->life_bound) cudaFree(c->life_bound);
    c->life_bound = nullptr; c->life_bound_words = 0;
    MB_CUDA(c, cudaMalloc(&c->life_bound, need * sizeof(unsigned long long)));
    c->life_bound_words = need;
  }
  if (ncta > c->life_nflags) {
    MB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->life_flags) cudaFree(c->life_flags);
    c->life_flags = nullptr; c->life_nflags = 0;
    MB_CUDA(c, cudaMalloc(&c->life_flags, ncta * sizeof(int)));
    c->life_nflags = ncta;
  }
  MB_CUDA(c, cudaMemsetAsync(c->life_flags, 0, ncta * sizeof(int), c->stream));
  LifeBitsParams P;
  P.h = h; P.w = w; P.wpr = wpr; P.rows_own = rows_own; P.T = T; P.wrap = p.border == MB_WRAP;
  P.ngen = (int)(n_steps - 1);
  P.bound = c->life_bound; P.flags = c->life_flags;
  static int configured[64] = {0};
  if (!configured[c->device & 63]) {
    MB_CUDA(c, cudaFuncSetAttribute(life_bits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured[c->device & 63] = 1;
  }
  const unsigned char *in = (const unsigned char *)d_b;
  unsigned char *out = (unsigned char *)d_a;
  void *args[] = {(void *)&P, (void *)&in, (void *)&out};
  // cooperative launch: all CTAs are co-resident, which the neighbour hand-shake relies on
  MB_CUDA(c, cudaLaunchCooperativeKernel((const void *)life_bits_kernel, dim3(ncta), dim3(256), args, smem, c->stream));
  c->launches++;
  c->last_kernel = "life_bits";
  *result = d_a;
  return MB_OK;
}

}  // namespace mb
