// Micro-benchmarks that decide the design of the update kernel: issue rate / latency of
// FFMA, mma.sync.m16n8k8.tf32 (legacy tensor path), IEEE vs fast div/sqrt, cluster.sync.
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int NACC>
__global__ void k_mma(float* out, long long* cyc, int iters) {
  float d[NACC][4];
  unsigned a[4] = {threadIdx.x, threadIdx.x + 1, 3, 4}, b[2] = {5, threadIdx.x};
#pragma unroll
  for (int i = 0; i < NACC; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) mma_tf32(d[i], a, b);
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// the access pattern of spo_warp_mma_3xtf32<1>: per k-step 6 mma on NCH independent accumulator chains
template <int NCH>
__global__ void k_mma_chain(float* out, long long* cyc, int iters) {
  float d[NCH][4];
  unsigned a[4] = {threadIdx.x, threadIdx.x + 1, 3, 4}, b[2] = {5, threadIdx.x};
#pragma unroll
  for (int i = 0; i < NCH; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 6; ++j) mma_tf32(d[j % NCH], a, b);
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int NACC>
__global__ void k_ffma(float* out, long long* cyc, int iters, float x, float y) {
  float d[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) d[i] = i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) d[i] = fmaf(d[i], x, y);
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
__global__ void k_div(float* out, long long* cyc, int iters, float x) {
  float d[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) d[i] = 1.5f + i + threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) d[i] = __fdiv_rn(x, __fadd_rn(__fsqrt_rn(d[i]), 1e-8f));
      else d[i] = __fdividef(x, __fsqrt_rn(d[i]) + 1e-8f);
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_cluster(long long* cyc, int iters) {
  cg::cluster_group c = cg::this_cluster();
  __shared__ float slot[2];
  c.sync();
  long long t0 = clock64();
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    if (threadIdx.x == 0) slot[it & 1] = it;
    c.sync();
    for (unsigned b = 0; b < c.num_blocks(); ++b) acc += *c.map_shared_rank(slot + (it & 1), b);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = (t1 - t0) + (acc < 0 ? 1 : 0);
  c.sync();
}

int main() {
  float* out; long long* cyc; long long h[8];
  cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 64 * sizeof(long long));
  const int iters = 2000;
  int threads[] = {128, 256, 512, 1024};
  for (int t : threads) {
    k_mma<8><<<1, t>>>(out, cyc, iters); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)h[0] / iters / 8;  // cycles per mma per warp
    printf("mma.m16n8k8.tf32 indep x8: threads=%4d  %.2f cyc/mma/warp -> %.0f FMA/clk/SM\n", t, per, 1024.0 * (t / 32) / per);
  }
  k_mma<1><<<1, 32>>>(out, cyc, iters); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("mma.m16n8k8.tf32 dependent chain latency: %.1f cyc\n", (double)h[0] / iters);
  k_mma<2><<<1, 32>>>(out, cyc, iters); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("mma 2 indep per warp, 1 warp: %.1f cyc per pair\n", (double)h[0] / iters);
  k_mma<4><<<1, 128>>>(out, cyc, iters); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("mma 4 indep, 4 warps (1/SMSP): %.2f cyc/mma/warp\n", (double)h[0] / iters / 4);
  for (int t : {256, 512}) {
    k_mma_chain<2><<<1, t>>>(out, cyc, iters); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("6 mma/k-step on 2 chains (3 deep), threads=%d: %.1f cyc per k-step -> 64^3 GEMM (8 k-steps) %.0f cyc\n", t, (double)h[0] / iters, 8.0 * h[0] / iters);
    k_mma_chain<6><<<1, t>>>(out, cyc, iters); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("6 mma/k-step on 6 chains (independent), threads=%d: %.1f cyc per k-step\n", t, (double)h[0] / iters);
    k_mma_chain<3><<<1, t>>>(out, cyc, iters); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("6 mma/k-step on 3 chains, threads=%d: %.1f cyc per k-step\n", t, (double)h[0] / iters);
  }
  for (int t : threads) {
    k_ffma<16><<<1, t>>>(out, cyc, iters, 1.0001f, 0.5f); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)h[0] / iters / 16;
    printf("ffma indep x16: threads=%4d  %.2f cyc/ffma/warp -> %.0f FMA/clk/SM\n", t, per, 32.0 * (t / 32) / per);
  }
  k_ffma<1><<<1, 32>>>(out, cyc, iters, 1.0001f, 0.5f); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("ffma dependent latency: %.1f cyc\n", (double)h[0] / iters);
  k_div<0><<<1, 256>>>(out, cyc, 500, 3.f); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("IEEE sqrt+div, 256 thr: %.1f cyc per (sqrt,add,div) per warp\n", (double)h[0] / 500 / 8);
  k_div<1><<<1, 256>>>(out, cyc, 500, 3.f); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("IEEE sqrt + fast div, 256 thr: %.1f cyc per warp\n", (double)h[0] / 500 / 8);
  for (int cs : {2, 3, 4, 6, 8}) {
    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(cs); cfg.blockDim = dim3(256);
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_cluster, cyc, 1000);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e != cudaSuccess || e2 != cudaSuccess) { printf("cluster=%d: launch failed: %s / %s\n", cs, cudaGetErrorString(e), cudaGetErrorString(e2)); cudaGetLastError(); continue; }
    cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("cluster=%d: %.0f cyc per (sync + DSMEM read of all slots)\n", cs, (double)h[0] / 1000);
  }
  return 0;
}
