// Bring-up test of the tcgen05 building block used by the full-batch forward kernel:
// D[128 x 64] (TMEM, fp32) = A[128 x 64] * B[64 x 64]^T with kind::tf32, operands in shared
// memory in the canonical K-major / no-swizzle ("interleave") layout, optional 3xTF32
// error compensation.  Verifies descriptor / instruction-descriptor encodings, TMEM
// alloc + tcgen05.ld, commit -> mbarrier.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, K = 64;

// canonical K-major no-swizzle: 8-row x 16-byte core matrices.
//   byte offset of element (r, k) = (r % 8) * 16 + (r / 8) * SBO + (k / 4) * LBO + (k % 4) * 4
// with the 16 K-chunks of an 8-row group stored back to back: LBO = 128 B, SBO = 16 * 128 = 2048 B.
constexpr uint32_t LBO = 128, SBO = 2048;
__host__ __device__ inline uint32_t canon_off(int r, int k) { return (r % 8) * 16 + (r / 8) * SBO + (k / 4) * LBO + (k % 4) * 4; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);        // start address
  d |= static_cast<uint64_t>((LBO >> 4) & 0x3FFF) << 16;    // leading byte offset (between K chunks)
  d |= static_cast<uint64_t>((SBO >> 4) & 0x3FFF) << 32;    // stride byte offset (between 8-row groups)
  d |= 1ull << 46;                                          // descriptor version (Blackwell)
  return d;                                                 // layout_type = 0 (SWIZZLE_NONE), base_offset = 0
}

__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
  uint32_t d = 0;
  d |= 1u << 4;                 // c_format = F32
  d |= 2u << 7;                 // a_format = TF32
  d |= 2u << 10;                // b_format = TF32
  d |= static_cast<uint32_t>(n >> 3) << 17;
  d |= static_cast<uint32_t>(m >> 4) << 24;
  return d;                     // a_major = b_major = 0 (K-major)
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum));
}

__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__global__ void __launch_bounds__(128) tc_kernel(const float* A, const float* B, float* D, int split) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* a_hi = reinterpret_cast<float*>(smem);                       // 128*64*4 = 32 KB
  float* a_lo = reinterpret_cast<float*>(smem + 32768);
  float* b_hi = reinterpret_cast<float*>(smem + 65536);               // 64*64*4 = 16 KB
  float* b_lo = reinterpret_cast<float*>(smem + 65536 + 16384);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < M * K; i += 128) {
    int r = i / K, k = i % K;
    float v = A[i];
    float hi = (split >= 2) ? rna_tf32(v) : __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a_hi) + canon_off(r, k)) = (split >= 2) ? hi : v;
    *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a_lo) + canon_off(r, k)) = (split >= 2) ? rna_tf32(v - hi) : v - hi;
  }
  for (int i = tid; i < N * K; i += 128) {
    int r = i / K, k = i % K;
    float v = B[i];
    float hi = (split >= 2) ? rna_tf32(v) : __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(b_hi) + canon_off(r, k)) = (split >= 2) ? hi : v;
    *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(b_lo) + canon_off(r, k)) = (split >= 2) ? rna_tf32(v - hi) : v - hi;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // generic-proxy smem writes -> visible to the async proxy (UMMA reads)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_d = tmem_base;

  if (warp == 0 && (tid & 31) == 0) {
    const uint32_t idesc = make_idesc(M, N);
    uint32_t accum = 0;
    // split: 0 = 1xTF32, 1 = 3xTF32 (truncating split), 2 = 3xTF32 (rounded split), 3 = 4xTF32 (rounded split)
    const int npass = split == 0 ? 1 : (split == 3 ? 4 : 3);
    for (int pass = npass - 1; pass >= 0; --pass) {   // small terms first
      const float* ap = (pass >= 2) ? a_lo : a_hi;
      const float* bp = (pass == 1 || pass == 3) ? b_lo : b_hi;
      for (int ks = 0; ks < K / 8; ++ks) {
        const uint64_t da = make_desc(smem_u32(ap) + ks * 2 * LBO);
        const uint64_t db = make_desc(smem_u32(bp) + ks * 2 * LBO);
        mma_tf32(tmem_d, da, db, idesc, accum);
        accum = 1;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // everyone waits for the MMAs
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  // epilogue: thread = row (warp w reads TMEM lanes [32w, 32w+32)), 64 columns in two x32 loads
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
  for (int j = 0; j < N; ++j) D[tid * N + j] = __uint_as_float(v[j]);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem_d));
}

int main() {
  float *hA = new float[M * K], *hB = new float[N * K], *hD = new float[M * N];
  srand(1);
  for (int i = 0; i < M * K; ++i) hA[i] = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  for (int i = 0; i < N * K; ++i) hB[i] = (rand() / (float)RAND_MAX - 0.5f) * 0.5f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, M * K * 4); cudaMalloc(&dB, N * K * 4); cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, hA, M * K * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, N * K * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int split = 0; split < 4; ++split) {
    cudaMemset(dD, 0, M * N * 4);
    tc_kernel<<<1, 128, 100 * 1024>>>(dA, dB, dD, split);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("split=%d CUDA error: %s\n", split, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost);
    double max_abs = 0, max_ref = 0;
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < N; ++j) {
        double s = 0;
        for (int k = 0; k < K; ++k) s += (double)hA[i * K + k] * hB[j * K + k];
        max_abs = fmax(max_abs, fabs(s - hD[i * N + j]));
        max_ref = fmax(max_ref, fabs(s));
      }
    printf("split=%d: max |err| = %.3e (max |ref| = %.3f, rel %.2e)  D[0][0]=%f D[5][7]=%f\n", split, max_abs, max_ref, max_abs / max_ref,
           hD[0], hD[5 * N + 7]);
  }
  return 0;
}
