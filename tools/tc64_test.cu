// Bring-up test for the next step of the update kernel: the 64x64x64 products of a
// minibatch step on tcgen05 with M = 64, operands in ONE blocked shared-memory layout
//     [row/8][col/4][row%8][col%4]   (128-byte core matrices; 16 KB per 64x64 fp32 tile)
// that is at the same time
//   * the canonical K-major  no-swizzle layout of (MN = row, K = col): SBO = 2048 B (row/8), LBO = 128 B (col/4), and
//   * the canonical MN-major no-swizzle layout of (MN = col, K = row): SBO = 128 B (col/4), LBO = 2048 B (row/8),
// so a tile written once by its producer feeds both the forward product (A = h[r][k], K-major) and the
// weight-gradient product (dW[j][k] = sum_r dz[r][j] * h[r][k]: both operands MN-major).
//
// Variants:  0: D = A * B^T, both K-major (forward layers, dh1's A)      1: D = P^T * Q, both MN-major (dW2, dW1)
//            2: A K-major, B MN-major (dh1 = dz2 * W2 with W2 stored [j][k])
// 3xTF32 with the truncating split (hi = the fp32 word itself, lo = x - trunc(x) as a second tile), small terms first.
// Accumulator: TMEM, M = 64 -> row m lives in datapath lane (m % 16) + 32 * (m / 16) (cute tmem_frg, M_MMA == 64).
// Prints max |err| against fp64 and the cycles from first issue to the commit's mbarrier completion.
// First run (last GPU seconds of round 1, profiles/r01_tc64_test.txt): variant 0 is CORRECT (max |err| 1.1e-6, which confirms the
// M = 64 datapath-lane mapping) and the 24-instruction product completes in 1378 cycles from first issue (57 per
// instruction against the 32-cycle floor; the mma.sync version of the same product takes ~2600).  Variants 1 and 2
// return zeros: the MN-major descriptor / instruction-descriptor setup below is still wrong -- first thing to fix next round
// (candidates: LBO/SBO roles for MN-major no-swizzle, the per-k-step start-address advance, a 128B-swizzled layout instead).
// Build with
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tc64_test tools/tc64_test.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int T = 64;   // every tile is 64 x 64
__host__ __device__ inline uint32_t blk_off(int row, int col) {   // byte offset in the blocked layout
  return ((row >> 3) * 16 + (col >> 2)) * 128 + (row & 7) * 16 + (col & 3) * 4;
}
constexpr uint32_t ROWGRP = 2048, COLCHK = 128;
// alternative blocked layout for the MN-major experiments: [col/4][row/8][row%8][col%4] (row groups of one column chunk contiguous)
__host__ __device__ inline uint32_t blk_off_alt(int row, int col) {
  return ((col >> 2) * 8 + (row >> 3)) * 128 + (row & 7) * 16 + (col & 3) * 4;
}
// 128B-swizzled layout: [col/32][row][col%32] (128-byte rows, 8 KB per 32-column atom), 16-byte chunk index XORed with row % 8.
// As K-major (MN = row, K = col): SBO = 1024 (8-row groups), a k-step of 8 advances the start address by 32 B inside the atom and
// by 8192 B to the next atom.  As MN-major (MN = col, K = row): LBO = 8192 (next 32-column atom), SBO = 1024, k-step = 1024.
__host__ __device__ inline uint32_t sw128_off(int row, int col) {
  uint32_t b = (col >> 5) * 8192 + row * 128 + (col & 31) * 4;
  return b ^ (((b >> 7) & 7) << 4);
}
struct MnCfg {   // descriptor parameters tried for an MN-major operand
  int alt_layout;            // 0: blk_off, 1: blk_off_alt, 2: sw128_off (then ALL operands use the swizzled layout, layout_type = 2)
  uint32_t lbo, sbo, kstep;  // bytes
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout_type = 0) {
  uint64_t d = static_cast<uint64_t>(layout_type & 7) << 61;   // 0 none, 1 128B_BASE32B, 2 128B, 4 64B, 6 32B
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;   // Blackwell descriptor version; layout_type 0 = no swizzle
  return d;
}

__device__ __forceinline__ uint32_t make_idesc(int m, int n, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                 // c_format = F32
  d |= 2u << 7;                 // a_format = TF32
  d |= 2u << 10;                // b_format = TF32
  d |= static_cast<uint32_t>(a_mn_major) << 15;
  d |= static_cast<uint32_t>(b_mn_major) << 16;
  d |= static_cast<uint32_t>(n >> 3) << 17;
  d |= static_cast<uint32_t>(m >> 4) << 24;
  return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum));
}

// X, Y: plain row-major [64][64] in global memory.  variant 0: D[m][n] = sum_k X[m][k] Y[n][k]
//                                                   variant 1: D[m][n] = sum_k X[k][m] Y[k][n]
//                                                   variant 2: D[m][n] = sum_k X[m][k] Y[k][n]
__global__ void __launch_bounds__(128) tc64_kernel(const float* X, const float* Y, float* D, int variant, long long* cycles, MnCfg mn) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* x_hi = smem;
  uint8_t* x_lo = smem + 16384;
  uint8_t* y_hi = smem + 32768;
  uint8_t* y_lo = smem + 49152;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < T * T; i += 128) {
    const int r = i / T, c = i % T;
    const float xv = X[i], yv = Y[i];
    const float xh = __uint_as_float(__float_as_uint(xv) & 0xFFFFE000u), yh = __uint_as_float(__float_as_uint(yv) & 0xFFFFE000u);
    const bool x_mn = (variant == 1), y_mn = (variant >= 1);
    const uint32_t xo = mn.alt_layout == 2 ? sw128_off(r, c) : ((x_mn && mn.alt_layout) ? blk_off_alt(r, c) : blk_off(r, c));
    const uint32_t yo = mn.alt_layout == 2 ? sw128_off(r, c) : ((y_mn && mn.alt_layout) ? blk_off_alt(r, c) : blk_off(r, c));
    *reinterpret_cast<float*>(x_hi + xo) = xv;        // the tensor core reads the top 19 bits
    *reinterpret_cast<float*>(x_lo + xo) = xv - xh;
    *reinterpret_cast<float*>(y_hi + yo) = yv;
    *reinterpret_cast<float*>(y_lo + yo) = yv - yh;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_d = tmem_base;

  long long t0 = 0;
  if (warp == 0 && lane == 0) {
    const int a_mn = (variant == 1), b_mn = (variant >= 1);
    const uint32_t idesc = make_idesc(T, T, a_mn, b_mn);
    // K-major operand (MN = row, K = col): LBO = column-chunk stride, SBO = row-group stride, a k-step of 8 = 2 chunks
    // MN-major operand (MN = col, K = row): LBO = row-group stride, SBO = column-chunk stride, a k-step of 8 = 1 row group
    const uint32_t a_lbo = a_mn ? mn.lbo : COLCHK, a_sbo = a_mn ? mn.sbo : ROWGRP, a_step = a_mn ? mn.kstep : 2 * COLCHK;
    const uint32_t b_lbo = b_mn ? mn.lbo : COLCHK, b_sbo = b_mn ? mn.sbo : ROWGRP, b_step = b_mn ? mn.kstep : 2 * COLCHK;
    const bool sw = mn.alt_layout == 2;
    uint32_t accum = 0;
    t0 = clock64();
    for (int pass = 2; pass >= 0; --pass) {   // lo*hi, hi*lo, hi*hi
      const uint8_t* ap = (pass == 2) ? x_lo : x_hi;
      const uint8_t* bp = (pass == 1) ? y_lo : y_hi;
      for (int ks = 0; ks < T / 8; ++ks) {
        uint64_t da, db;
        if (!sw) {
          da = make_desc(smem_u32(ap) + ks * a_step, a_lbo, a_sbo);
          db = make_desc(smem_u32(bp) + ks * b_step, b_lbo, b_sbo);
        } else {
          // swizzled K-major: 32 B per k-step inside a 128 B atom row, next atom after 4 k-steps; LBO unused (1), SBO = 1024
          const uint32_t koff = (ks & 3) * 32 + (ks >> 2) * 8192;
          da = a_mn ? make_desc(smem_u32(ap) + ks * mn.kstep, mn.lbo, mn.sbo, 2) : make_desc(smem_u32(ap) + koff, 16, 1024, 2);
          db = b_mn ? make_desc(smem_u32(bp) + ks * mn.kstep, mn.lbo, mn.sbo, 2) : make_desc(smem_u32(bp) + koff, 16, 1024, 2);
        }
        mma_tf32(tmem_d, da, db, idesc, accum);
        accum = 1;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
  }
  if (warp == 0 && lane == 0) cycles[variant] = clock64() - t0;
  asm volatile("tcgen05.fence::after_thread_sync;");
  // epilogue: warp w owns datapath lanes [32w, 32w+32); with M = 64 its lanes 0..15 hold rows 16w .. 16w+15
  uint32_t v[64];
  const uint32_t taddr = tmem_d + (static_cast<uint32_t>(warp * 32) << 16);
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[c * 32 + 0]), "=r"(v[c * 32 + 1]), "=r"(v[c * 32 + 2]), "=r"(v[c * 32 + 3]), "=r"(v[c * 32 + 4]), "=r"(v[c * 32 + 5]),
          "=r"(v[c * 32 + 6]), "=r"(v[c * 32 + 7]), "=r"(v[c * 32 + 8]), "=r"(v[c * 32 + 9]), "=r"(v[c * 32 + 10]), "=r"(v[c * 32 + 11]),
          "=r"(v[c * 32 + 12]), "=r"(v[c * 32 + 13]), "=r"(v[c * 32 + 14]), "=r"(v[c * 32 + 15]), "=r"(v[c * 32 + 16]), "=r"(v[c * 32 + 17]),
          "=r"(v[c * 32 + 18]), "=r"(v[c * 32 + 19]), "=r"(v[c * 32 + 20]), "=r"(v[c * 32 + 21]), "=r"(v[c * 32 + 22]), "=r"(v[c * 32 + 23]),
          "=r"(v[c * 32 + 24]), "=r"(v[c * 32 + 25]), "=r"(v[c * 32 + 26]), "=r"(v[c * 32 + 27]), "=r"(v[c * 32 + 28]), "=r"(v[c * 32 + 29]),
          "=r"(v[c * 32 + 30]), "=r"(v[c * 32 + 31])
        : "r"(taddr + c * 32));
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  if (lane < 16)
    for (int j = 0; j < T; ++j) D[(warp * 16 + lane) * T + j] = __uint_as_float(v[j]);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem_d));
}

int main() {
  float *hX = new float[T * T], *hY = new float[T * T], *hD = new float[T * T];
  srand(1);
  for (int i = 0; i < T * T; ++i) { hX[i] = (rand() / (float)RAND_MAX - 0.5f) * 2.f; hY[i] = (rand() / (float)RAND_MAX - 0.5f) * 0.5f; }
  float *dX, *dY, *dD;
  long long* dC;
  cudaMalloc(&dX, T * T * 4); cudaMalloc(&dY, T * T * 4); cudaMalloc(&dD, T * T * 4); cudaMalloc(&dC, 3 * sizeof(long long));
  cudaMemcpy(dX, hX, T * T * 4, cudaMemcpyHostToDevice); cudaMemcpy(dY, hY, T * T * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(tc64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  // candidates for the MN-major descriptors (variant 0 ignores them).  #0 is what cute's make_umma_desc<Major::MN> derives for
  // the no-swizzle layout ((T,1,m),(8,k)):((1,T,SBO),(1T,LBO)); it returned zeros in the first run, hence the scan.
  const MnCfg cands[] = {{0, ROWGRP, COLCHK, ROWGRP}, {0, COLCHK, ROWGRP, ROWGRP}, {1, COLCHK, 8 * COLCHK, COLCHK}, {1, 8 * COLCHK, COLCHK, COLCHK},
                         {0, ROWGRP, COLCHK, 2 * ROWGRP}, {1, COLCHK, 8 * COLCHK, 2 * COLCHK},
                         // 128B-swizzled tiles (cute: MN-major B128 = ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO))); also run for variant 0 as a check
                         // of the swizzled K-major path
                         {2, 8192, 1024, 1024}, {2, 1024, 8192, 1024}};
  const int ncand = sizeof(cands) / sizeof(cands[0]);
  for (int variant = 0; variant < 3; ++variant)
    for (int ci = 0; ci < ncand; ++ci) {
      const MnCfg mn = cands[ci];
      if (variant == 0 && ci != 0 && mn.alt_layout != 2) continue;   // K-major only: the no-swizzle layout once, the swizzled one once
      if (variant == 0 && ci == ncand - 1) continue;
      cudaMemset(dD, 0, T * T * 4);
      for (int rep = 0; rep < 2; ++rep) tc64_kernel<<<1, 128, 66 * 1024>>>(dX, dY, dD, variant, dC, mn);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("variant %d cand %d CUDA error: %s\n", variant, ci, cudaGetErrorString(e)); return 1; }
      long long cyc[3];
      cudaMemcpy(hD, dD, T * T * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(cyc, dC, sizeof(cyc), cudaMemcpyDeviceToHost);
      double max_abs = 0, max_ref = 0, sum_abs = 0;
      for (int m = 0; m < T; ++m)
        for (int n = 0; n < T; ++n) {
          double sacc = 0;
          for (int k = 0; k < T; ++k) {
            const double a = (variant == 1) ? hX[k * T + m] : hX[m * T + k];
            const double b = (variant == 0) ? hY[n * T + k] : hY[k * T + n];
            sacc += a * b;
          }
          max_abs = fmax(max_abs, fabs(sacc - hD[m * T + n]));
          max_ref = fmax(max_ref, fabs(sacc));
          sum_abs += fabs(hD[m * T + n]);
        }
      printf("variant %d cand %d (alt_layout %d lbo %u sbo %u kstep %u): max |err| = %.3e (max |ref| = %.3f, mean |D| = %.3e)   24 mma: %lld cycles\n",
             variant, ci, mn.alt_layout, mn.lbo, mn.sbo, mn.kstep, max_abs, max_ref, sum_abs / (T * T), cyc[variant]);
    }
  return 0;
}
