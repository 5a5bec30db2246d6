// Validates the warp-level 3xTF32 mma.sync tile routine (csrc/spo_mma.cuh) against a CPU
// reference for the operand orientations the update kernel uses.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>
#include "../safe-policy-optimization_b200/csrc/spo_mma.cuh"

// C[64 x 64] = A(m,k) * B(k,n), strides given; each of 8 warps computes a 32 x 16 patch
__global__ void k_test(const float* A, int a_sm, int a_sk, const float* B, int b_sk, int b_sn, int K, float* C) {
  extern __shared__ float sm[];
  float* As = sm;
  float* Bs = sm + 64 * 72;
  for (int i = threadIdx.x; i < 64 * 72; i += blockDim.x) { As[i] = A[i]; Bs[i] = B[i]; }
  __syncthreads();
  float acc[2][2][4];
  spo_mma_zero<2>(acc);
  const int warp = threadIdx.x >> 5;
  spo_warp_mma_3xtf32<2>(acc, As, a_sm, a_sk, Bs, b_sk, b_sn, (warp & 1) * 32, (warp >> 1) * 16, K);
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  for (int mt = 0; mt < 2; ++mt)
    for (int nt = 0; nt < 2; ++nt) {
      const int r = (warp & 1) * 32 + mt * 16 + g, c = (warp >> 1) * 16 + nt * 8 + 2 * t;
      C[r * 64 + c] = acc[mt][nt][0]; C[r * 64 + c + 1] = acc[mt][nt][1];
      C[(r + 8) * 64 + c] = acc[mt][nt][2]; C[(r + 8) * 64 + c + 1] = acc[mt][nt][3];
    }
}

int main() {
  const int LD = 72, K = 64;
  float *hA = new float[64 * LD], *hB = new float[64 * LD], *hC = new float[64 * 64];
  srand(3);
  for (int i = 0; i < 64 * LD; ++i) { hA[i] = (rand() / (float)RAND_MAX - 0.5f) * 2.f; hB[i] = (rand() / (float)RAND_MAX - 0.5f); }
  float *dA, *dB, *dC;
  cudaMalloc(&dA, 64 * LD * 4); cudaMalloc(&dB, 64 * LD * 4); cudaMalloc(&dC, 64 * 64 * 4);
  cudaMemcpy(dA, hA, 64 * LD * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, 64 * LD * 4, cudaMemcpyHostToDevice);
  // orientation cases: (a_sm, a_sk, b_sk, b_sn)
  int cases[4][4] = {{LD, 1, 1, LD}, {1, LD, LD, 1}, {LD, 1, LD, 1}, {1, LD, 1, LD}};
  for (auto& c : cases) {
    k_test<<<1, 256, 2 * 64 * 72 * 4>>>(dA, c[0], c[1], dB, c[2], c[3], K, dC);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hC, dC, 64 * 64 * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < 64; ++m)
      for (int n = 0; n < 64; ++n) {
        double s = 0;
        for (int k = 0; k < K; ++k) s += (double)hA[m * c[0] + k * c[1]] * hB[k * c[2] + n * c[3]];
        maxerr = fmax(maxerr, fabs(s - hC[m * 64 + n])); maxref = fmax(maxref, fabs(s));
      }
    printf("strides A(%d,%d) B(%d,%d): max|err| %.3e  max|ref| %.3f\n", c[0], c[1], c[2], c[3], maxerr, maxref);
  }
  return 0;
}
