// The two per-step transforms between env.step and the policy forward:
//
//  * spo_obs_normalize -- SafeNormalizeObservation (safepo/common/wrappers.py:42-49), i.e.
//    gymnasium's NormalizeObservation / RunningMeanStd (third-party, not under the reference
//    checkout; algorithm restated in oracle/envio.py):
//        batch_mean, batch_var (population) over the n envs        (float64)
//        delta = batch_mean - mean;  tot = count + n
//        mean' = mean + delta * n / tot
//        var'  = (var*count + batch_var*n + delta^2 * count * n / tot) / tot
//        out   = (obs - mean') / sqrt(var' + eps)                   (eps = 1e-8)
//  * spo_action_rescale -- SafeRescaleAction(env, -1, 1) (safepo/common/env.py:62,76), gymnasium's
//    RescaleAction: clip to [min_action, max_action], map affinely onto [low, high], clip.
//
// One CTA owns 32 observation columns: coalesced row-major reads with 8 row lanes, two-pass
// variance in fp64 like numpy's np.var, the new statistics written by the column's first row
// lane, then the same CTA normalises its columns (second read is L2-hot).  n x D is a few
// hundred KB at most: the kernel is launch-latency bound, one launch per env step.
#include "spo_common.cuh"

namespace {

constexpr int NC = 32;   // columns per CTA
constexpr int NR = 8;    // row lanes

__global__ void __launch_bounds__(NC * NR) spo_obs_normalize_kernel(const float* obs, int n, int D, double* mean, double* var,
                                                                    double count, double* count_out, int update, double eps,
                                                                    float* out /* may alias obs */) {
  __shared__ double red[NR][NC];
  __shared__ double s_mean[NC], s_var[NC];
  const int cx = threadIdx.x, ry = threadIdx.y;
  const int col = blockIdx.x * NC + cx;
  const bool cv = col < D;
  if (update) {
    double s = 0.0;
    if (cv)
      for (int r = ry; r < n; r += NR) s += static_cast<double>(obs[static_cast<size_t>(r) * D + col]);
    red[ry][cx] = s;
    __syncthreads();
    double bm = 0.0;
#pragma unroll
    for (int i = 0; i < NR; ++i) bm += red[i][cx];
    bm /= static_cast<double>(n);
    __syncthreads();
    double q = 0.0;
    if (cv)
      for (int r = ry; r < n; r += NR) {
        const double d = static_cast<double>(obs[static_cast<size_t>(r) * D + col]) - bm;
        q += d * d;
      }
    red[ry][cx] = q;
    __syncthreads();
    if (ry == 0 && cv) {
      double bv = 0.0;
#pragma unroll
      for (int i = 0; i < NR; ++i) bv += red[i][cx];
      bv /= static_cast<double>(n);
      const double bc = static_cast<double>(n), tot = count + bc;
      const double m0 = mean[col], v0 = var[col];
      const double delta = bm - m0;
      const double m1 = m0 + delta * bc / tot;
      const double M2 = v0 * count + bv * bc + delta * delta * count * bc / tot;
      const double v1 = M2 / tot;
      mean[col] = m1;
      var[col] = v1;
      s_mean[cx] = m1;
      s_var[cx] = v1;
    }
    if (blockIdx.x == 0 && cx == 0 && ry == 0 && count_out) *count_out = count + static_cast<double>(n);
  } else if (ry == 0 && cv) {
    s_mean[cx] = mean[col];
    s_var[cx] = var[col];
  }
  __syncthreads();
  if (cv && out) {
    const double m = s_mean[cx], sd = sqrt(s_var[cx] + eps);
    for (int r = ry; r < n; r += NR) {
      const size_t i = static_cast<size_t>(r) * D + col;
      // (obs - mean) / sqrt(var + eps) in float64, rounded once to fp32 (the reference converts the
      // float64 wrapper output with torch.as_tensor(..., dtype=float32))
      out[i] = static_cast<float>((static_cast<double>(obs[i]) - m) / sd);
    }
  }
}

__global__ void spo_action_rescale_kernel(const float* __restrict__ act, int n, int A, const float* __restrict__ low,
                                          const float* __restrict__ high, float min_a, float max_a, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * A) return;
  const int j = i % A;
  const float lo = low[j], hi = high[j];
  float a = fminf(fmaxf(act[i], min_a), max_a);
  a = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), __fdiv_rn(__fsub_rn(a, min_a), __fsub_rn(max_a, min_a))));
  out[i] = fminf(fmaxf(a, lo), hi);
}

}  // namespace

extern "C" int spo_obs_normalize(const float* obs, int n, int obs_dim, double* mean, double* var, double count, double* count_out,
                                 int update, double eps, float* out, void* stream) {
  SPO_REQUIRE(obs && mean && var, SPO_ERR_INVALID_ARG, "spo_obs_normalize: null obs / mean / var");
  SPO_REQUIRE(n >= 1 && obs_dim >= 1, SPO_ERR_INVALID_ARG, "spo_obs_normalize: n=%d obs_dim=%d", n, obs_dim);
  SPO_REQUIRE(update || out, SPO_ERR_INVALID_ARG, "spo_obs_normalize: nothing to do (update=0 and out=NULL)");
  SPO_REQUIRE(count > 0.0, SPO_ERR_INVALID_ARG, "spo_obs_normalize: count must be positive (RunningMeanStd starts at 1e-4)");
  dim3 block(NC, NR);
  spo_obs_normalize_kernel<<<(obs_dim + NC - 1) / NC, block, 0, static_cast<cudaStream_t>(stream)>>>(obs, n, obs_dim, mean, var, count, count_out,
                                                                                                   update, eps, out);
  SPO_CUDA_TRY(cudaGetLastError());
  return SPO_OK;
}

extern "C" int spo_action_rescale(const float* act, int n, int act_dim, const float* low, const float* high, float min_action, float max_action,
                                  float* out, void* stream) {
  SPO_REQUIRE(act && low && high && out, SPO_ERR_INVALID_ARG, "spo_action_rescale: null argument");
  SPO_REQUIRE(n >= 1 && act_dim >= 1 && max_action > min_action, SPO_ERR_INVALID_ARG, "spo_action_rescale: n=%d act_dim=%d range [%g, %g]", n,
              act_dim, min_action, max_action);
  const int total = n * act_dim;
  spo_action_rescale_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(act, n, act_dim, low, high, min_action, max_action, out);
  SPO_CUDA_TRY(cudaGetLastError());
  return SPO_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Row G2 (multi-agent path, BASELINE config 5): SeparatedReplayBuffer.compute_returns / compute_cost_returns
// (safepo/common/buffer.py:356-384) -- masked GAE on PopArt-de-normalised value predictions, fp32 throughout
// like the reference, time-major [T(+1)][N] so one thread per env reads coalesced rows:
//     D(x)  = x * sqrt(var) + mean                                  (popart.py:128)
//     delta = ((r_t + (gamma * D(v_{t+1})) * m_{t+1}) - D(v_t)
//     gae   = delta + ((gamma*lambda) * m_{t+1}) * gae
//     ret_t = gae + D(v_t)
// every product and sum rounded separately, in the reference's association (torch evaluates the python expression
// left to right, scalars cast to fp32), so the sequential kernel reproduces it bit for bit.
// STATUS: written at the end of round 1 after the GPU budget was spent -- compiled, covered by an xfail-tolerant GPU test
// (tests/test_zz_pending_gpu.py), not yet run on hardware.
namespace {
__global__ void spo_gae_masked_kernel(const float* __restrict__ rewards, const float* __restrict__ vpred, const float* __restrict__ masks,
                                      float mean, float sqrt_var, float gamma, float gamma_lambda, float* __restrict__ returns, int N, int T) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float gae = 0.f;
  float d_next = __fadd_rn(__fmul_rn(vpred[static_cast<size_t>(T) * N + n], sqrt_var), mean);   // D(v_T): the bootstrap row
  for (int t = T - 1; t >= 0; --t) {
    const size_t i = static_cast<size_t>(t) * N + n;
    const float m = masks[i + N];
    const float d_cur = __fadd_rn(__fmul_rn(vpred[i], sqrt_var), mean);
    const float delta = __fsub_rn(__fadd_rn(rewards[i], __fmul_rn(__fmul_rn(gamma, d_next), m)), d_cur);
    gae = __fadd_rn(delta, __fmul_rn(__fmul_rn(gamma_lambda, m), gae));
    returns[i] = __fadd_rn(gae, d_cur);
    d_next = d_cur;
  }
}
}  // namespace

extern "C" int spo_gae_masked(const float* rewards, const float* value_preds, const float* masks, float popart_mean, float popart_sqrt_var,
                              float gamma, double gamma_lambda, float* returns, int N, int T, void* stream) {
  SPO_REQUIRE(rewards && value_preds && masks && returns, SPO_ERR_INVALID_ARG, "spo_gae_masked: null argument");
  SPO_REQUIRE(N >= 1 && T >= 1, SPO_ERR_INVALID_ARG, "spo_gae_masked: N=%d T=%d", N, T);
  spo_gae_masked_kernel<<<(N + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(rewards, value_preds, masks, popart_mean, popart_sqrt_var,
                                                                                     gamma, static_cast<float>(gamma_lambda), returns, N, T);
  SPO_CUDA_TRY(cudaGetLastError());
  return SPO_OK;
}
