// The minibatch update loop as one persistent launch per pass.
//
// Reference: safepo/single_agent/ppo_lag.py:297-336 (PPO-Lag), focops.py:309-357
// (FOCOPS), cpo.py:543-571 / trpo_lag.py:466-494 (critic regression).  Per minibatch:
// forward of the three nets, losses, backward, critic L2 term, ONE joint grad-norm clip
// over all three nets (ppo_lag.py:325), three Adam steps.
//
// GEMMs: every 64x64xK product of the step runs on the tensor pipe as warp-level
// mma.sync.m16n8k8 TF32 with the 3xTF32 split in registers (csrc/spo_mma.cuh) -- the FFMA
// register-tile version of r01 was bound by shared-memory bandwidth at 4.1-4.8 k cycles per
// GEMM (profiles/r01_update_phase_cycles.md).  Accumulator fragments of the dW products
// are the gradients; their owner threads also hold the Adam moments in registers.
//
// Mapping: a thread-block cluster of 4 CTAs (the 4th only joins the barriers: a cluster of
// 4 synchronises faster than one of 3 on B200), one net per CTA -- actor / reward critic / cost
// critic, 512 threads each.  A CTA keeps its net's weights, the Adam moments and all
// activations of the 64-row tile in shared memory / registers for the whole pass; per step it
// reads the 64 minibatch rows (staged in registers a step ahead, their indices two steps
// ahead) and exchanges one float through distributed shared memory for the joint gradient
// norm.  Weights and moments touch HBM once per launch.
//
// The chain of steps is strictly sequential (each step needs the weights of the previous
// one), so this kernel is latency-bound by construction: what is optimised is
// microseconds per step, not bandwidth.  Every phase between two barriers is spread over all
// 16 warps (loss rows as (row, action dim), reductions as 8-way row splits + shuffles):
// with 4 warps per scheduler an instruction executed by every thread costs 4 issue cycles,
// and a phase run by one warp stalls the other 15 (profiles/r01_update_phase_cycles.md).
#include <cooperative_groups.h>
#include <stdlib.h>
#include "spo_common.cuh"
#include "spo_mma.cuh"

namespace cg = cooperative_groups;

// Optional phase timers (build with -DSPO_PHASE_TIMERS): thread 0 of every CTA accumulates
// clock64() deltas per phase of the step; read back with spo_debug_phase_cycles().
#ifdef SPO_PHASE_TIMERS_V1
__device__ unsigned long long g_phase_cycles[4][16];
__device__ long long g_trace[4][16][24];   // arrival time of every warp at every mark during step 50
#define TRACE_MARK(idx)                                                   \
  do {                                                                    \
    if (lane == 0 && step_idx == 50) g_trace[rank & 3][wid][idx] = clock64(); \
  } while (0)
#define PHASE_MARK(idx)                                                   \
  do {                                                                    \
    TRACE_MARK(idx);                                                      \
    if (tid == 0) {                                                       \
      const long long now__ = clock64();                                  \
      sm_phase__[idx] += static_cast<unsigned long long>(now__ - phase_t__); /* no global traffic inside the step */ \
      phase_t__ = now__;                                                  \
    }                                                                     \
  } while (0)
#else
#define PHASE_MARK(idx) do { } while (0)
#define TRACE_MARK(idx) do { } while (0)
#endif

namespace {

constexpr int AUXW = 28;       // per-row side data: act[8] | logp adv tgt _ | old_mean[8] | old_std[8]
constexpr int AUX_LOGP = 8, AUX_ADV = 9, AUX_TGT = 10, AUX_OMEAN = 12, AUX_OSTD = 20;
constexpr float kLogSqrt2Pi = 0.91893853320467274178f;

struct UpdArgsV1 {
  float *params, *adam_m, *adam_v;
  int* adam_t;
  spo_batch data;
  const int64_t* perm;
  int64_t perm_len;
  int batch, kind, D, A;
  int actor_only;   // CUP projection stage: the critic CTAs only join the barriers
  spo_hparams hp;
  spo_update_ctrl* ctrl;
  spo_comm comm;   // world <= 1: single GPU
};

// system-scope accesses for the cross-GPU gradient exchange (peer memory over NVLink)
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_relaxed_sys_f4(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_relaxed_sys_f(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem))), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// One Adam step on a scalar in torch's _multi_tensor_adam op order:
//   m = lerp(m, g, 1-b1) (fused mul-add);  v = v*b2 + ((1-b2)*g)*g;
//   denom = sqrt(v)/sqrt(bc2) + eps;  p = p + (step_size*m)/denom,  step_size = -lr/bc1.
// sqrt and the division use the SFU approximations (sqrt.approx / div.approx,
// <= 2 ulp), 1/sqrt(bc2) is a precomputed factor: IEEE-exact versions cost ~60 issue slots per parameter (profiles/r01) for
// differences far below the 1e-5 parity bar.
struct AdamK {
  float w1, b2, w2, ibc2s, eps, ss;  // w1 = 1-b1, w2 = 1-b2, ibc2s = 1/sqrt(1-b2^t)
};
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float adam_update(float p, float g, float& m, float& v, const AdamK& k) {
  m = fmaf(k.w1, __fsub_rn(g, m), m);
  v = __fadd_rn(__fmul_rn(v, k.b2), __fmul_rn(__fmul_rn(k.w2, g), g));
  const float denom = fmaf(sqrt_approx(v), k.ibc2s, k.eps);
  return __fadd_rn(p, __fdividef(__fmul_rn(k.ss, m), denom));
}

// small parameters of a net in the order b1[64] b2[64] w3[O*64] b3[O] log_std[A(actor)]
struct SmallMap {
  int O, A_ls;  // A_ls = A for the actor, 0 for critics
  __device__ int count() const { return 2 * SPO_HID + O * SPO_HID + O + A_ls; }
  // global offset (within packed buffer) of small param i
  __device__ int goff(const SpoNetOff& o, int i) const {
    if (i < SPO_HID) return o.b1 + i;
    i -= SPO_HID;
    if (i < SPO_HID) return o.b2 + i;
    i -= SPO_HID;
    if (i < O * SPO_HID) return o.w3 + i;
    i -= O * SPO_HID;
    if (i < O) return o.b3 + i;
    return o.log_std + (i - O);
  }
};

constexpr int UT = 512;   // threads per CTA: 16 warps, 4 per scheduler -- every phase of the step is a short dependent
                          // chain, so latency hiding (not issue width) sets the pace (profiles/r01_update_phase_cycles.md)
constexpr int FE = 8;     // accumulator-fragment elements per thread per 64x64 product (16 x 16 patch per warp)

// Every shared-memory extent is a compile-time constant: the observation tile / W1 image are padded to
// KX = 64 * NT1 input columns (zeros beyond obs_dim), leading dimension KX + 4 (== 4 mod 8: rows g = 0..7 of
// an mma fragment fall in 8 different bank groups), the output layer and the small-parameter slots are
// sized for SPO_MAX_ACT.  Addresses are then immediates off one base -- with runtime extents the
// compiler rematerialised pointer arithmetic inside the step (18 % of the issued instructions, profiles/r01).
__host__ __device__ constexpr int upd_ldx(int nt1) { return 64 * nt1 + 4; }
constexpr int SPN = (2 * SPO_HID + SPO_MAX_ACT * SPO_HID + 2 * SPO_MAX_ACT + 3) & ~3;   // small-parameter slots

// element e = nt*4 + c of the 16 x 16 warp patch of a 64-wide output owned by thread tid
__device__ __forceinline__ void frag_rc(int tid, int e8, int col_base, int& row, int& col) {
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int nt = e8 >> 2, e = e8 & 3;
  row = (warp & 3) * 16 + g + ((e >> 1) << 3);
  col = col_base + (warp >> 2) * 16 + nt * 8 + 2 * t + (e & 1);
}

template <int NT1>
__global__ void __launch_bounds__(UT, 1) spo_update_v1_kernel(const UpdArgsV1 a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ int comm_dead;   // a peer never showed up: stop waiting (ctrl->stop = 2 tells the host)
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  if (threadIdx.x == 0) comm_dead = 0;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  spo_update_ctrl* ctrl = a.ctrl;
  if (*reinterpret_cast<volatile int*>(&ctrl->stop)) return;  // whole cluster takes this branch together

  const int D = a.D, A = a.A;
  constexpr int KX = 64 * NT1, ldx = upd_ldx(NT1);
  const bool idle = rank >= 3;
  const int net = idle ? 2 : static_cast<int>(rank);
  const bool is_actor = (net == 0) && !idle;
  const bool active = !idle && !(is_actor && a.kind == SPO_LOSS_CRITIC_ONLY) && !(!is_actor && a.actor_only);
  const SpoNetOff off = spo_net_off(D, A, net);
  const int O = off.out;
  const SmallMap sm{O, is_actor ? A : 0};
  const int SP = sm.count();
  const int mb = (wid & 3) * 16, nb = (wid >> 2) * 16;   // 16 x 16 warp patch of every 64x64 product
  const int g8 = lane >> 2, t4 = lane & 3;

  // ---- shared memory carve-up (weights in nn.Linear orientation [out][in]) ----
  float* p = smem;
  // minibatch row indices of tile q in slot q & 1: requested a step ahead of the rows they address, so the
  // gather never waits on a dependent global load
  int64_t* idxbuf = reinterpret_cast<int64_t*>(p); p += 2 * 2 * SPO_ROWS;   // [2][64] int64
  float* lsc = p; p += 4 * SPO_MAX_ACT;           // per action dim: std, 1/var, log(std), spare (refreshed every step)
  float* adk = p; p += 8;                         // Adam scalars of the current step (one thread does the fp64 math)
  float* w1 = p;  p += SPO_HID * ldx;
  float* b1 = p;  p += SPO_HID;
  float* w2 = p;  p += SPO_HID * SPO_LDH;
  float* b2 = p;  p += SPO_HID;
  // every CTA of the cluster carves the SAME layout (actor-sized output layer): the peers read
  // xchg through distributed shared memory at their own offset of it
  float* w3 = p;  p += SPO_MAX_ACT * SPO_HID;
  float* b3 = p;  p += SPO_MAX_ACT;
  float* log_std = p; p += 8;
  float* msmall = p;  p += SPN;
  float* vsmall = p;  p += SPN;
  float* gsmall = p;  p += SPN;
  float* x = p;   p += SPO_ROWS * ldx;      // observation tile (the next one is staged in registers)
  float* aux = p; p += SPO_ROWS * AUXW;     // per-row side data
  float* h1 = p;  p += SPO_ROWS * SPO_LDH;
  float* h2 = p;  p += SPO_ROWS * SPO_LDH;   // becomes dz1 during backward
  float* dz2 = p; p += SPO_ROWS * SPO_LDH;
  float* y = p;   p += SPO_ROWS * SPO_MAX_ACT;
  float* dy = p;  p += SPO_ROWS * SPO_MAX_ACT;
  float* dls = p; p += SPO_ROWS * SPO_MAX_ACT;   // per-row d loss / d log_std
  float* red = p; p += 64;                       // block-reduction scratch
  float* xchg = p; p += 4;                       // [parity] CTA grad sumsq, read by peers through DSMEM
  float* mv1b = nullptr;                         // moments of the second W1 column block (obs_dim > 64): thread-private slots
  if (NT1 > 1) { mv1b = p; p += 2 * FE * UT; }
  float* dz1 = h2;

  const int tps = (a.batch + SPO_ROWS - 1) / SPO_ROWS;                    // tiles per step
  const int64_t n_steps = (a.perm_len + a.batch - 1) / a.batch;
  const int64_t n_tiles = n_steps * tps;

  // ---- one-time loads ----
  if (!idle) {
    for (int i = tid; i < SPO_HID * ldx; i += UT) {
      const int j = i / ldx, k = i - j * ldx;
      w1[i] = (k < D) ? __ldg(a.params + off.w1 + j * D + k) : 0.f;
    }
    for (int i = tid; i < SPO_HID * SPO_LDH; i += UT) {
      const int j = i / SPO_LDH, k = i - j * SPO_LDH;
      w2[i] = (k < SPO_HID) ? __ldg(a.params + off.w2 + j * SPO_HID + k) : 0.f;
    }
    for (int i = tid; i < SPO_HID; i += UT) { b1[i] = __ldg(a.params + off.b1 + i); b2[i] = __ldg(a.params + off.b2 + i); }
    for (int i = tid; i < O * SPO_HID; i += UT) w3[i] = __ldg(a.params + off.w3 + i);
    for (int i = tid; i < O; i += UT) b3[i] = __ldg(a.params + off.b3 + i);
    for (int i = tid; i < SP; i += UT) {
      msmall[i] = a.adam_m[sm.goff(off, i)];
      vsmall[i] = a.adam_v[sm.goff(off, i)];
    }
    if (is_actor && tid < A) log_std[tid] = a.params[off.log_std + tid];
    for (int i = tid; i < SPO_ROWS * ldx; i += UT) x[i] = 0.f;
    for (int i = tid; i < SPO_ROWS * AUXW; i += UT) aux[i] = 0.f;
  }
  // Adam moments of this thread's fragment elements: W2 and the first 64 input columns of W1 in registers
  float mW2[FE], vW2[FE], mW1[FE], vW1[FE];
  if (active) {
#pragma unroll
    for (int e = 0; e < FE; ++e) {
      int j, k;
      frag_rc(tid, e, 0, j, k);
      mW2[e] = a.adam_m[off.w2 + j * SPO_HID + k];
      vW2[e] = a.adam_v[off.w2 + j * SPO_HID + k];
      const bool ok = k < D;
      mW1[e] = ok ? a.adam_m[off.w1 + j * D + k] : 0.f;
      vW1[e] = ok ? a.adam_v[off.w1 + j * D + k] : 0.f;
      if (NT1 > 1) {
        const bool ok2 = (k + 64) < D;
        mv1b[e * UT + tid] = ok2 ? a.adam_m[off.w1 + j * D + k + 64] : 0.f;
        mv1b[(FE + e) * UT + tid] = ok2 ? a.adam_v[off.w1 + j * D + k + 64] : 0.f;
      }
    }
  }
  const int t0 = idle ? 0 : a.adam_t[net];
  double b1pow = pow(static_cast<double>(a.hp.beta1), static_cast<double>(t0));
  double b2pow = pow(static_cast<double>(a.hp.beta2), static_cast<double>(t0));
  const float lr = (net == 0) ? a.hp.lr_actor : (net == 1 ? a.hp.lr_reward : a.hp.lr_cost);
  const float extra_sumsq = (is_actor && a.kind == SPO_LOSS_CRITIC_ONLY) ? ctrl->extra_sumsq : 0.f;
  const float vcoef = (net == 1) ? a.hp.value_coef : 1.f;
  const float reg = is_actor ? 0.f : __fmul_rn(vcoef, __fmul_rn(a.hp.critic_l2, 2.f));

  // The next tile is staged in registers: its rows are requested with plain loads at the end of the current
  // step (between the arrive and the wait of the cluster barrier) and stored to shared memory at the top of the
  // next one, Adam in between hides the latency.  (cp.async gathers cost 1.5 k cycles of issue per tile for
  // the rows plus 1.7 k for the side data -- LDGSTS issues at ~50 cycles per warp instruction here,
  // profiles/r01_update_phase_cycles.md.)
  // Item it of this thread: obs_dim % 4 == 0: float4 chunk i = tid + it*UT -> (row, chunk) = (i / (D/4), i % (D/4)),
  // decoded once; otherwise scalar element i -> (i / D, i % D).
  constexpr int PF_MAX = (NT1 == 1) ? 2 : 4;
  const bool vec_rows = (D & 3) == 0;
  int pf_rc[PF_MAX];
  int pf_n = 0;
#pragma unroll
  for (int it = 0; it < PF_MAX; ++it) {
    const int i = tid + it * UT, per_row = D >> 2;
    pf_rc[it] = 0xFF;
    if (vec_rows && i < SPO_ROWS * per_row) { pf_rc[it] = (i / per_row) | ((i % per_row) << 8); pf_n = it + 1; }
  }
  float xr[4 * PF_MAX];   // staged observation values
  float auxr[4];          // staged side data: column c = q8s + 8*i of row r8s
#pragma unroll
  for (int i = 0; i < 4 * PF_MAX; ++i) xr[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) auxr[i] = 0.f;
  const int r8s = tid >> 3, q8s = tid & 7;
  const int aux_per = !is_actor ? 1 : A + 2 + (a.kind == SPO_LOSS_FOCOPS ? 2 * A : 0);   // <= 26 columns
  // shared-memory slot of side-data column c
  auto aux_slot = [&](int c) {
    if (!is_actor) return AUX_TGT;
    if (c < A) return c;
    if (c == A) return AUX_LOGP;
    if (c == A + 1) return AUX_ADV;
    if (c < 2 * A + 2) return AUX_OMEAN + (c - A - 2);
    return AUX_OSTD + (c - 2 * A - 2);
  };

  // global source of side-data column c (row g at src + g * (A or 1))
  auto aux_by_row = [&](int c) { return is_actor && (c < A || c >= A + 2); };
  auto aux_src = [&](int c) -> const float* {
    if (!is_actor) return (net == 1) ? a.data.target_r : a.data.target_c;
    if (c < A) return a.data.act + c;
    if (c == A) return a.data.logp;
    if (c == A + 1) return a.data.adv;
    if (c < 2 * A + 2) return a.data.old_mean + (c - A - 2);
    return a.data.old_std + (c - 2 * A - 2);
  };
  const float* aux_src0 = aux_src(q8s < aux_per ? q8s : 0);   // column q8s: decoded once
  const int aux_mul0 = aux_by_row(q8s) ? A : 1;
  const int aux_slot0 = aux_slot(q8s < aux_per ? q8s : 0);

  int64_t step_idx = 0;
#ifdef SPO_PHASE_TIMERS_V1
  __shared__ unsigned long long sm_phase__[16];
  if (tid < 16) sm_phase__[tid] = 0ull;
#endif
  // tile (step, sub) after n more tiles -- no 64-bit division on the per-step path
  auto tile_after = [&](int64_t step, int sub, int n, int64_t& step_o, int& sub_o) {
    step_o = step; sub_o = sub;
    for (int i = 0; i < n; ++i)
      if (++sub_o == tps) { sub_o = 0; ++step_o; }
  };
  // request the rows of tile q = (step, sub): global -> registers (rows beyond the valid range stage zeros)
  auto load_next = [&](int64_t q, int64_t step, int sub) {
    if (!active || q >= n_tiles) return;
    int64_t rs = a.perm_len - step * a.batch;
    if (rs > a.batch) rs = a.batch;
    int rows = static_cast<int>(rs) - sub * SPO_ROWS;
    rows = rows < 0 ? 0 : (rows > SPO_ROWS ? SPO_ROWS : rows);
    const int64_t* ridx = idxbuf + (q & 1) * SPO_ROWS;
    if (vec_rows) {
#pragma unroll
      for (int it = 0; it < PF_MAX; ++it) {
        const int r = pf_rc[it] & 0xFF, c = pf_rc[it] >> 8;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (it < pf_n && r < rows) v = __ldg(reinterpret_cast<const float4*>(a.data.obs + ridx[r] * D + 4 * c));
        xr[4 * it] = v.x; xr[4 * it + 1] = v.y; xr[4 * it + 2] = v.z; xr[4 * it + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int it = 0; it < 4 * PF_MAX; ++it) {
        const int i = tid + it * UT, r = i / D, c = i - r * D;
        xr[it] = (r < rows) ? __ldg(a.data.obs + ridx[r] * D + c) : 0.f;
      }
    }
    const bool rv = r8s < rows;
    const int64_t g = rv ? ridx[r8s] : 0;
    auxr[0] = (rv && q8s < aux_per) ? __ldg(aux_src0 + g * aux_mul0) : 0.f;
    if (aux_per > 8) {   // wide action spaces only
#pragma unroll
      for (int i = 1; i < 4; ++i) {
        const int c = q8s + 8 * i;
        float v = 0.f;
        if (rv && c < aux_per) v = __ldg(aux_src(c) + g * (aux_by_row(c) ? A : 1));
        auxr[i] = v;
      }
    }
  };
  // registers -> the tile buffers
  auto store_next = [&]() {
    if (!active) return;
    if (vec_rows) {
#pragma unroll
      for (int it = 0; it < PF_MAX; ++it) {
        const int r = pf_rc[it] & 0xFF, c = pf_rc[it] >> 8;
        if (it < pf_n) *reinterpret_cast<float4*>(x + r * ldx + 4 * c) = make_float4(xr[4 * it], xr[4 * it + 1], xr[4 * it + 2], xr[4 * it + 3]);
      }
    } else {
#pragma unroll
      for (int it = 0; it < 4 * PF_MAX; ++it) {
        const int i = tid + it * UT, r = i / D, c = i - r * D;
        if (r < SPO_ROWS) x[r * ldx + c] = xr[it];
      }
    }
    if (q8s < aux_per) aux[r8s * AUXW + aux_slot0] = auxr[0];
    if (aux_per > 8) {
#pragma unroll
      for (int i = 1; i < 4; ++i) {
        const int c = q8s + 8 * i;
        if (c < aux_per) aux[r8s * AUXW + aux_slot(c)] = auxr[i];
      }
    }
  };
  // row indices of tile q -> idxbuf slot q & 1 (lands with the cp.async group it is committed in)
  auto fetch_idx = [&](int64_t q, int64_t step, int sub) {
    if (!active || q >= n_tiles) return;
    const int64_t first = step * a.batch + sub * SPO_ROWS;
    int64_t rs = a.perm_len - first;
    if (rs > a.batch - sub * SPO_ROWS) rs = a.batch - sub * SPO_ROWS;
    if (tid < SPO_ROWS && tid < rs) cp_async8(idxbuf + (q & 1) * SPO_ROWS + tid, a.perm + first + tid);
  };

#ifdef SPO_PHASE_TIMERS_V1
  long long phase_t__ = clock64();
#endif
  // hidden layer: out[r][j] = tanh(b[j] + sum_k in[r][k] * W[j][k]) on the tensor pipe.  The 64-wide case
  // (second layer always; first layer when obs_dim pads to 64) gets compile-time strides: immediate
  // offsets instead of per-load address arithmetic (2.7 k -> 1.7 k cycles per product)
  auto hidden = [&](const float* in, int ldin, int K, const float* W, int ldw, const float* bias, float* out) {
    float acc[1][2][4];
    spo_mma_zero<1>(acc);
    spo_warp_mma_3xtf32<1>(acc, in, ldin, 1, W, 1, ldw, mb, nb, K);   // all extents are constants after inlining
    PHASE_MARK(11);  // (sub) hidden-layer GEMM only, as seen by warp 0
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int r = mb + g8, j = nb + nt * 8 + 2 * t4;
      const float2 bb = *reinterpret_cast<const float2*>(bias + j);
      *reinterpret_cast<float2*>(out + r * SPO_LDH + j) = make_float2(spo_tanh_fast(acc[0][nt][0] + bb.x), spo_tanh_fast(acc[0][nt][1] + bb.y));
      *reinterpret_cast<float2*>(out + (r + 8) * SPO_LDH + j) = make_float2(spo_tanh_fast(acc[0][nt][2] + bb.x), spo_tanh_fast(acc[0][nt][3] + bb.y));
    }
  };
  // column sums over the 64 rows of a [64][SPO_LDH] tile, all 512 threads: thread (c = tid >> 3, q = tid & 7)
  // adds rows q, q+8, ... (bank = 4q + c: conflict-free), three shuffles finish the sum
  auto colsum_into = [&](const float* buf, float* dst) {
    const int c = tid >> 3, q8 = tid & 7;
    float s = 0.f;
#pragma unroll
    for (int rr = 0; rr < SPO_ROWS / 8; ++rr) s += buf[(rr * 8 + q8) * SPO_LDH + c];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (q8 == 0) dst[c] += s;
  };

  __syncthreads();
  {
    int64_t st; int sb;
    fetch_idx(0, 0, 0);
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
    load_next(0, 0, 0);
    tile_after(0, 0, 1, st, sb);
    fetch_idx(1, st, sb);
    cp_async_commit();
  }

  // gradient accumulators = accumulator fragments of the dW products (persist across the tiles of a step)
  float gW2[FE], gW1[NT1][FE];
#pragma unroll
  for (int e = 0; e < FE; ++e) {
    gW2[e] = 0.f;
#pragma unroll
    for (int i = 0; i < NT1; ++i) gW1[i][e] = 0.f;
  }
  for (int i = tid; i < SP; i += UT) gsmall[i] = 0.f;
  double acc_loss = 0.0;        // thread 0: sum over steps of this net's logged loss
  float step_loss = 0.f;        // thread 0: loss numerator of the current step (sum over its tiles)
  float step_aux0 = 0.f, step_aux1 = 0.f;  // FOCOPS: sum(ratio*adv), sum(mask)

  int64_t step = 0;   // tile q = (step, sub)
  int sub = 0;
  auto next_tile = [&]() { if (++sub == tps) { sub = 0; ++step; } };
  for (int64_t q = 0; q < n_tiles; ++q, next_tile()) {
    int64_t rs64 = a.perm_len - step * a.batch;
    if (rs64 > a.batch) rs64 = a.batch;
    const int rows_step = static_cast<int>(rs64);
    int rows = rows_step - sub * SPO_ROWS;
    rows = rows < 0 ? 0 : (rows > SPO_ROWS ? SPO_ROWS : rows);
    const bool last_tile = (sub == tps - 1);
    const float inv_b = __fdiv_rn(1.f, static_cast<float>(rows_step));
    // stage tile q+1 (its indices landed a step ago) and request the indices of tile q+2
    auto stage_next = [&]() {
      int64_t st; int sb;
      tile_after(step, sub, 1, st, sb);
      load_next(q + 1, st, sb);
      tile_after(st, sb, 1, st, sb);
      fetch_idx(q + 2, st, sb);
      cp_async_commit();
    };

    store_next();      // tile q: registers -> shared memory (every warp left tile q-1 barriers ago)
    TRACE_MARK(22);
    __syncthreads();   // tile q in place; Adam's weight writes visible
    TRACE_MARK(15);
    if (is_actor && tid >= UT - 32 && tid - (UT - 32) < A) {
      // row-independent pieces of the Gaussian log-density (log_std changed in the last Adam step)
      const int j = tid - (UT - 32);
      const float sd = expf(log_std[j]);
      lsc[4 * j + 0] = sd;
      lsc[4 * j + 1] = __fdiv_rn(1.f, __fmul_rn(sd, sd));
      lsc[4 * j + 2] = logf(sd);
    }
    PHASE_MARK(0);   // top of the step: stage-in + barrier

    if (active) {
      // ---------------- forward ----------------
      hidden(x, ldx, KX, w1, ldx, b1, h1);
      __syncthreads();
      PHASE_MARK(1);
      hidden(h1, SPO_LDH, SPO_HID, w2, SPO_LDH, b2, h2);
      __syncthreads();
      PHASE_MARK(2);
      // output layer: thread (row r = tid >> 3, eighth q8 of k): h2 chunks q8 and q8+8 stay in registers for all
      // outputs (the tile is read from shared memory once); three shuffles finish each dot product
      const int r8 = tid >> 3, q8 = tid & 7;
      {
        const float4 ha = *reinterpret_cast<const float4*>(h2 + r8 * SPO_LDH + 4 * q8);
        const float4 hb = *reinterpret_cast<const float4*>(h2 + r8 * SPO_LDH + 32 + 4 * q8);
        for (int o = 0; o < O; ++o) {
          const float4 wa = *reinterpret_cast<const float4*>(w3 + o * SPO_HID + 4 * q8);
          const float4 wb = *reinterpret_cast<const float4*>(w3 + o * SPO_HID + 32 + 4 * q8);
          float sacc = (fmaf(ha.x, wa.x, ha.y * wa.y) + fmaf(ha.z, wa.z, ha.w * wa.w)) +
                       (fmaf(hb.x, wb.x, hb.y * wb.y) + fmaf(hb.z, wb.z, hb.w * wb.w));
          sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
          sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
          sacc += __shfl_xor_sync(0xffffffffu, sacc, 4);
          if (q8 == 0) y[r8 * SPO_MAX_ACT + o] = sacc + b3[o];
        }
      }
      TRACE_MARK(17);
      __syncthreads();
      if (tid == UT - 32 && last_tile) {
        // Adam scalars of this step, computed once per CTA by a thread that has no loss row; only this
        // thread tracks the beta powers (fp64, like torch's Python floats)
        b1pow *= static_cast<double>(a.hp.beta1);
        b2pow *= static_cast<double>(a.hp.beta2);
        adk[0] = static_cast<float>(1.0 - static_cast<double>(a.hp.beta1));
        adk[1] = a.hp.beta2;
        adk[2] = static_cast<float>(1.0 - static_cast<double>(a.hp.beta2));
        // the root and the quotients in fp32 (<= 1 ulp from torch's fp64-then-round)
        adk[3] = __fdiv_rn(1.f, sqrtf(static_cast<float>(1.0 - b2pow)));
        adk[4] = a.hp.adam_eps;
        adk[5] = -__fdiv_rn(lr, static_cast<float>(1.0 - b1pow));
      }

      // ---------------- loss and d loss / d output: thread (row r8, action dim j = q8) ----------------
      float part0 = 0.f, part1 = 0.f, part2 = 0.f;
      {
        const int r = r8, j = q8;
        const bool valid = r < rows;
        const float* ax = aux + r * AUXW;
        if (!is_actor) {
          if (j == 0) {
            const float dv = __fsub_rn(y[r * SPO_MAX_ACT], ax[AUX_TGT]);
            part0 = valid ? __fmul_rn(dv, dv) : 0.f;
            dy[r * SPO_MAX_ACT] = valid ? __fmul_rn(__fmul_rn(__fmul_rn(2.f, dv), inv_b), vcoef) : 0.f;
          }
        } else {
          const bool jv = j < A;
          float term = 0.f, klj = 0.f, dmu_lp = 0.f, dls_lp = 0.f, dmu_kl = 0.f, dls_kl = 0.f;
          if (jv) {
            const float mean = y[r * SPO_MAX_ACT + j];
            const float std = lsc[4 * j], inv_var = lsc[4 * j + 1];
            const float diff = __fsub_rn(ax[j], mean);
            const float d2 = __fmul_rn(diff, diff);
            const float q2 = __fmul_rn(d2, inv_var);           // (a - mu)^2 / var
            term = __fsub_rn(__fsub_rn(__fmul_rn(-0.5f, q2), lsc[4 * j + 2]), kLogSqrt2Pi);
            dmu_lp = __fmul_rn(diff, inv_var);
            dls_lp = __fsub_rn(q2, 1.f);
            if (a.kind == SPO_LOSS_FOCOPS) {
              // KL(new || old), torch _kl_normal_normal(p=new, q=old)
              // padded rows carry zeros: keep their (discarded) arithmetic finite
              const float os = valid ? ax[AUX_OSTD + j] : 1.f, om = ax[AUX_OMEAN + j];
              const float ios = __fdiv_rn(1.f, os);
              const float sr = __fmul_rn(std, ios);
              const float vr = __fmul_rn(sr, sr);
              const float dm = __fmul_rn(__fsub_rn(mean, om), ios);
              const float t1 = __fmul_rn(dm, dm);
              klj = __fmul_rn(0.5f, __fsub_rn(__fsub_rn(__fadd_rn(vr, t1), 1.f), logf(vr)));
              dmu_kl = __fmul_rn(dm, ios);
              dls_kl = __fsub_rn(vr, 1.f);
            }
          }
          // sums over the action dims: the 8 lanes of a row (lanes j >= A hold zeros)
          float lp = term, kl = klj;
          lp += __shfl_xor_sync(0xffffffffu, lp, 1);
          lp += __shfl_xor_sync(0xffffffffu, lp, 2);
          lp += __shfl_xor_sync(0xffffffffu, lp, 4);
          if (a.kind == SPO_LOSS_FOCOPS) {
            kl += __shfl_xor_sync(0xffffffffu, kl, 1);
            kl += __shfl_xor_sync(0xffffffffu, kl, 2);
            kl += __shfl_xor_sync(0xffffffffu, kl, 4);
          }
          const float ratio = expf(__fsub_rn(lp, ax[AUX_LOGP]));
          const float adv = ax[AUX_ADV];
          if (a.kind == SPO_LOSS_PPO_CLIP) {
            const float s1 = __fmul_rn(ratio, adv);
            const float s2 = __fmul_rn(fminf(fmaxf(ratio, a.hp.clip_lo), a.hp.clip_hi), adv);
            if (j == 0) part0 = valid ? -fminf(s1, s2) : 0.f;
            // d(-mean(min))/d logp = -(1/B) * adv * ratio where the unclipped branch is active
            const float gl = (valid && s1 <= s2) ? -__fmul_rn(__fmul_rn(adv, ratio), inv_b) : 0.f;
            if (jv) {
              dy[r * SPO_MAX_ACT + j] = __fmul_rn(gl, dmu_lp);
              dls[r * SPO_MAX_ACT + j] = __fmul_rn(gl, dls_lp);
            }
          } else {
            // FOCOPS needs mean(mask) over the minibatch before gradients can be formed:
            // stash per-row pieces, finish after the block reduction below.
            const float mask = (valid && kl <= a.hp.focops_kl) ? 1.f : 0.f;
            if (j == 0) {
              part0 = valid ? __fmul_rn(kl, mask) : 0.f;
              part1 = valid ? __fmul_rn(ratio, adv) : 0.f;
              part2 = mask;
            }
            if (jv) {
              // first term: (1/B) * mask * d kl ; second term scaled later by mean(mask)
              const float gl = valid ? __fmul_rn(__fmul_rn(adv, ratio), inv_b) : 0.f;
              dy[r * SPO_MAX_ACT + j] = __fmul_rn(__fmul_rn(mask, inv_b), dmu_kl);
              dls[r * SPO_MAX_ACT + j] = __fmul_rn(__fmul_rn(mask, inv_b), dls_kl);
              y[r * SPO_MAX_ACT + j] = __fmul_rn(gl, dmu_lp);
              // d logp / d log_std piece: into the (now free) old-mean slot of the current aux row
              const_cast<float*>(ax)[AUX_OMEAN + j] = __fmul_rn(gl, dls_lp);
            }
          }
        }
        part0 = spo_warp_sum(part0);
        if (is_actor && a.kind == SPO_LOSS_FOCOPS) { part1 = spo_warp_sum(part1); part2 = spo_warp_sum(part2); }
        if (lane == 0) { red[wid * 4 + 0] = part0; red[wid * 4 + 1] = part1; red[wid * 4 + 2] = part2; }
      }
      TRACE_MARK(18);
      __syncthreads();
      if (tid == 0) {
        float l0 = 0.f, l1 = 0.f, l2 = 0.f;
#pragma unroll
        for (int w = 0; w < UT / 32; ++w) { l0 += red[w * 4]; l1 += red[w * 4 + 1]; l2 += red[w * 4 + 2]; }
        step_loss += l0;
        step_aux0 += l1;
        step_aux1 += l2;
      }
      if (is_actor && a.kind == SPO_LOSS_FOCOPS) {
        // This formulation needs the whole minibatch in one tile (batch <= 64): mean(mask)
        // and the per-row pieces are combined here.  (focops.py uses batch 64.)
        float msum = 0.f;
#pragma unroll
        for (int w = 0; w < UT / 32; ++w) msum += red[w * 4 + 2];
        const float mbar = __fmul_rn(msum, inv_b);
        const float c2 = -__fmul_rn(__fdiv_rn(1.f, a.hp.focops_lam), mbar);
        if (q8 < A) {
          const float* ax = aux + r8 * AUXW;
          dy[r8 * SPO_MAX_ACT + q8] = __fadd_rn(dy[r8 * SPO_MAX_ACT + q8], __fmul_rn(c2, y[r8 * SPO_MAX_ACT + q8]));
          dls[r8 * SPO_MAX_ACT + q8] = __fadd_rn(dls[r8 * SPO_MAX_ACT + q8], __fmul_rn(c2, ax[AUX_OMEAN + q8]));
        }
        __syncthreads();
      }

      PHASE_MARK(3);   // output layer + loss rows
      // ---------------- backward ----------------
      // (a) small grads of the output layer: dW3[o][k] = sum_r dy[r][o] h2[r][k], thread (k = tid >> 3, q8): rows
      //     q8, q8+8, ... with the h2 column in registers for all outputs; three shuffles finish each sum.
      //     db3[o] / dlog_std[j] (column sums of dy / dls): the first four warps, same row split.
      {
        const int k = r8;
        float hv[SPO_ROWS / 8];
#pragma unroll
        for (int rr = 0; rr < SPO_ROWS / 8; ++rr) hv[rr] = h2[(rr * 8 + q8) * SPO_LDH + k];
        for (int o = 0; o < O; ++o) {
          float sa = 0.f;
#pragma unroll
          for (int rr = 0; rr < SPO_ROWS / 8; ++rr) sa = fmaf(dy[(rr * 8 + q8) * SPO_MAX_ACT + o], hv[rr], sa);
          sa += __shfl_xor_sync(0xffffffffu, sa, 1);
          sa += __shfl_xor_sync(0xffffffffu, sa, 2);
          sa += __shfl_xor_sync(0xffffffffu, sa, 4);
          if (q8 == 0) gsmall[2 * SPO_HID + o * SPO_HID + k] += sa;
        }
        if (tid < 128) {
          const int c = tid >> 3, col = c & 7;          // c < 8: dy column (db3), else dls column (dlog_std)
          const bool need = (c < 8) ? (col < O) : (col < sm.A_ls);
          const float* src = (c < 8) ? dy : dls;
          float sa = 0.f;
          if (need) {
#pragma unroll
            for (int rr = 0; rr < SPO_ROWS / 8; ++rr) sa += src[(rr * 8 + q8) * SPO_MAX_ACT + col];
          }
          sa += __shfl_xor_sync(0xffffffffu, sa, 1);
          sa += __shfl_xor_sync(0xffffffffu, sa, 2);
          sa += __shfl_xor_sync(0xffffffffu, sa, 4);
          if (need && q8 == 0) gsmall[2 * SPO_HID + O * SPO_HID + ((c < 8) ? col : O + col)] += sa;
        }
      }
      TRACE_MARK(19);
      // (b) dz2[r][k] = (sum_o dy[r][o] * w3[o][k]) * (1 - h2[r][k]^2)
      {
        const int r0 = (tid >> 4) * 2, kk = (tid & 15) * 4;
#pragma unroll
        for (int ri = 0; ri < 2; ++ri) {
          const int r = r0 + ri;
          float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int o = 0; o < O; ++o) {
            const float d = dy[r * SPO_MAX_ACT + o];
            const float4 wv = *reinterpret_cast<const float4*>(w3 + o * SPO_HID + kk);
            s.x = fmaf(d, wv.x, s.x); s.y = fmaf(d, wv.y, s.y); s.z = fmaf(d, wv.z, s.z); s.w = fmaf(d, wv.w, s.w);
          }
          const float4 h = *reinterpret_cast<const float4*>(h2 + r * SPO_LDH + kk);
          s.x *= (1.f - h.x * h.x); s.y *= (1.f - h.y * h.y); s.z *= (1.f - h.z * h.z); s.w *= (1.f - h.w * h.w);
          *reinterpret_cast<float4*>(dz2 + r * SPO_LDH + kk) = s;
        }
      }
      __syncthreads();
      PHASE_MARK(4);   // small grads + dz2
      // (c) dW2[j][k] += sum_r dz2[r][j] * h1[r][k];  db2[j] += sum_r dz2[r][j]
      spo_warp_mma_3xtf32<1>(reinterpret_cast<float (&)[1][2][4]>(gW2), dz2, 1, SPO_LDH, h1, SPO_LDH, 1, mb, nb, SPO_ROWS);
      PHASE_MARK(12);  // (sub) dW2 GEMM
      colsum_into(dz2, gsmall + SPO_HID);
      PHASE_MARK(13);  // (sub) db2 column sums
      // (d) dz1[r][k] = (sum_j dz2[r][j] * W2[j][k]) * (1 - h1[r][k]^2)   -> overwrites h2
      {
        float acc[1][2][4];
        spo_mma_zero<1>(acc);
        spo_warp_mma_3xtf32<1>(acc, dz2, SPO_LDH, 1, w2, SPO_LDH, 1, mb, nb, SPO_HID);
        PHASE_MARK(14);  // (sub) dh1 GEMM
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const int r = mb + g8, k = nb + nt * 8 + 2 * t4;
          const float2 ha = *reinterpret_cast<const float2*>(h1 + r * SPO_LDH + k);
          const float2 hb = *reinterpret_cast<const float2*>(h1 + (r + 8) * SPO_LDH + k);
          *reinterpret_cast<float2*>(dz1 + r * SPO_LDH + k) =
              make_float2(acc[0][nt][0] * (1.f - ha.x * ha.x), acc[0][nt][1] * (1.f - ha.y * ha.y));
          *reinterpret_cast<float2*>(dz1 + (r + 8) * SPO_LDH + k) =
              make_float2(acc[0][nt][2] * (1.f - hb.x * hb.x), acc[0][nt][3] * (1.f - hb.y * hb.y));
        }
      }
      __syncthreads();
      PHASE_MARK(5);   // dW2 + dh1
      // (e) dW1[j][k] += sum_r dz1[r][j] * x[r][k];  db1[j] += sum_r dz1[r][j]
#pragma unroll
      for (int i = 0; i < NT1; ++i)
        spo_warp_mma_3xtf32<1>(reinterpret_cast<float (&)[1][2][4]>(gW1[i]), dz1, 1, SPO_LDH, x, ldx, 1, mb, i * 64 + nb, SPO_ROWS);
      colsum_into(dz1, gsmall);
    }  // active

    PHASE_MARK(6);   // dW1
    cp_async_wait_all();   // indices of tile q+1 (requested a step ago); the barrier below publishes them
    if (!last_tile) {      // next tile of the same step accumulates into the same gradients
      __syncthreads();
      stage_next();
      continue;
    }

    // ---------------- cross-GPU gradient sum (data-parallel ranks), in rank order ----------------
    // Push protocol: every rank stores its gradient straight into each peer's staging slot
    // [parity][source rank][net] over NVLink (posted remote stores), fences, then raises the
    // peer's flag [source rank][net] = seq.  A receiver only polls and reads its OWN memory.
    // Slot reuse is safe: a slot of parity p is rewritten at step s+2, after this rank saw every
    // peer's flag of step s+1, which a peer raises only after it finished reading step s.
    if (a.comm.world > 1 && active) {
      __syncthreads();  // gsmall complete
      constexpr int Q = FE / 4;   // float4s per tile per thread
      const int world = a.comm.world, me = a.comm.rank;
      const unsigned seq = static_cast<unsigned>(a.comm.seq_base + static_cast<unsigned long long>(step_idx) + 1ull);
      const size_t slot = static_cast<size_t>(UT) * FE * (1 + NT1) + spo_pad4(2 * SPO_HID + A * SPO_HID + 2 * A);
      const size_t par_off = static_cast<size_t>(seq & 1u) * world * 3 * slot;
      for (int r = 0; r < world; ++r) {
        if (r == me) continue;
        float4* dst = reinterpret_cast<float4*>(a.comm.grad_bufs[r] + par_off + (static_cast<size_t>(me) * 3 + net) * slot);
#pragma unroll
        for (int i = 0; i < Q; ++i) dst[i * UT + tid] = make_float4(gW2[4 * i], gW2[4 * i + 1], gW2[4 * i + 2], gW2[4 * i + 3]);
#pragma unroll
        for (int t = 0; t < NT1; ++t)
#pragma unroll
          for (int i = 0; i < Q; ++i)
            dst[(Q + t * Q + i) * UT + tid] = make_float4(gW1[t][4 * i], gW1[t][4 * i + 1], gW1[t][4 * i + 2], gW1[t][4 * i + 3]);
        float* dst_small = reinterpret_cast<float*>(dst + (Q + Q * NT1) * UT);
        for (int i = tid; i < SP; i += UT) dst_small[i] = gsmall[i];
      }
      __threadfence_system();
      __syncthreads();
      if (tid < world && tid != me) st_release_sys(a.comm.flags[tid] + me * 3 + net, seq);
      if (tid < world && tid != me && !comm_dead) {
        const unsigned limit = a.comm.spin_limit ? a.comm.spin_limit : 400000000u;
        const unsigned* f = a.comm.flags[me] + tid * 3 + net;     // local memory
        unsigned polls = 0;
        while (static_cast<int>(ld_acquire_sys(f) - seq) < 0) {
          if (++polls > limit) { comm_dead = 1; atomicExch(&ctrl->stop, 2); break; }
        }
      }
      __syncthreads();
      float sW2[FE], sW1[NT1][FE];
#pragma unroll
      for (int e = 0; e < FE; ++e) {
        sW2[e] = 0.f;
#pragma unroll
        for (int t = 0; t < NT1; ++t) sW1[t][e] = 0.f;
      }
      for (int r = 0; r < world; ++r) {
        if (r == me) {
#pragma unroll
          for (int e = 0; e < FE; ++e) {
            sW2[e] += gW2[e];
#pragma unroll
            for (int t = 0; t < NT1; ++t) sW1[t][e] += gW1[t][e];
          }
        } else {
          const float4* src = reinterpret_cast<const float4*>(a.comm.grad_bufs[me] + par_off + (static_cast<size_t>(r) * 3 + net) * slot);
          float4 v2[Q], v1[NT1][Q];
#pragma unroll
          for (int i = 0; i < Q; ++i) v2[i] = ld_relaxed_sys_f4(src + i * UT + tid);
#pragma unroll
          for (int t = 0; t < NT1; ++t)
#pragma unroll
            for (int i = 0; i < Q; ++i) v1[t][i] = ld_relaxed_sys_f4(src + (Q + t * Q + i) * UT + tid);
#pragma unroll
          for (int i = 0; i < Q; ++i) {
            sW2[4 * i] += v2[i].x; sW2[4 * i + 1] += v2[i].y; sW2[4 * i + 2] += v2[i].z; sW2[4 * i + 3] += v2[i].w;
#pragma unroll
            for (int t = 0; t < NT1; ++t) {
              sW1[t][4 * i] += v1[t][i].x; sW1[t][4 * i + 1] += v1[t][i].y; sW1[t][4 * i + 2] += v1[t][i].z; sW1[t][4 * i + 3] += v1[t][i].w;
            }
          }
        }
      }
      const float inv_w = __fdiv_rn(1.f, static_cast<float>(world));
#pragma unroll
      for (int e = 0; e < FE; ++e) {
        gW2[e] = __fmul_rn(sW2[e], inv_w);
#pragma unroll
        for (int t = 0; t < NT1; ++t) gW1[t][e] = __fmul_rn(sW1[t][e], inv_w);
      }
      for (int i = tid; i < SP; i += UT) {
        float sm_ = 0.f;
        for (int r = 0; r < world; ++r) {
          if (r == me) sm_ += gsmall[i];
          else sm_ += ld_relaxed_sys_f(reinterpret_cast<const float*>(reinterpret_cast<const float4*>(
                          a.comm.grad_bufs[me] + par_off + (static_cast<size_t>(r) * 3 + net) * slot) + (Q + Q * NT1) * UT) + i);
        }
        gsmall[i] = __fmul_rn(sm_, inv_w);
      }
    }

    PHASE_MARK(7);   // cross-GPU gradient exchange
    // ---------------- joint gradient norm (cluster-wide), clip, Adam ----------------
    float ss = 0.f, th2 = 0.f;
    if (active) {
      __syncthreads();  // gsmall complete
      if (is_actor) {   // no regulariser, no logged L2 term: nothing to read back
        // (dW1 columns >= obs_dim are exact zeros: the padded observation columns are)
#pragma unroll
        for (int e = 0; e < FE; ++e) {
          ss = fmaf(gW2[e], gW2[e], ss);
#pragma unroll
          for (int i = 0; i < NT1; ++i) ss = fmaf(gW1[i][e], gW1[i][e], ss);
        }
        for (int i = tid; i < SP; i += UT) ss = fmaf(gsmall[i], gsmall[i], ss);
      } else {
#pragma unroll
      for (int e = 0; e < FE; ++e) {
        int j, k;
        frag_rc(tid, e, 0, j, k);
        const float th = w2[j * SPO_LDH + k];
        const float g = fmaf(reg, th, gW2[e]);
        gW2[e] = g;
        ss = fmaf(g, g, ss);
        th2 = fmaf(th, th, th2);
#pragma unroll
        for (int i = 0; i < NT1; ++i) {
          const int kk = k + 64 * i;
          if (kk < D) {
            const float t1 = w1[j * ldx + kk];
            const float g1 = fmaf(reg, t1, gW1[i][e]);
            gW1[i][e] = g1;
            ss = fmaf(g1, g1, ss);
            th2 = fmaf(t1, t1, th2);
          } else {
            gW1[i][e] = 0.f;
          }
        }
      }
      for (int i = tid; i < SP; i += UT) {
        float th;
        if (i < SPO_HID) th = b1[i];
        else if (i < 2 * SPO_HID) th = b2[i - SPO_HID];
        else if (i < 2 * SPO_HID + O * SPO_HID) th = w3[i - 2 * SPO_HID];
        else if (i < 2 * SPO_HID + O * SPO_HID + O) th = b3[i - 2 * SPO_HID - O * SPO_HID];
        else th = log_std[i - 2 * SPO_HID - O * SPO_HID - O];
        const float g = fmaf(reg, th, gsmall[i]);
        gsmall[i] = g;
        ss = fmaf(g, g, ss);
        th2 = fmaf(th, th, th2);
      }
      }
      ss = spo_warp_sum(ss);
      if (!is_actor) th2 = spo_warp_sum(th2);
      if (lane == 0) { red[16 + wid] = ss; red[32 + wid] = th2; }
      TRACE_MARK(20);
      __syncthreads();
    }
    const int par = static_cast<int>(step_idx & 1);
    if (wid == 0) {
      float s = 0.f, t2 = 0.f;
      if (active && lane < UT / 32) { s = red[16 + lane]; t2 = red[32 + lane]; }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        t2 += __shfl_xor_sync(0xffffffffu, t2, o);
      }
      s += extra_sumsq;
      if (tid == 0) {
      xchg[par] = idle ? 0.f : s;
      // logged loss of this step (ppo_lag.py:330-336): critics include the L2 term
      if (active) {
        float L;
        if (!is_actor) L = fmaf(a.hp.critic_l2, t2, __fmul_rn(step_loss, inv_b));
        else if (a.kind == SPO_LOSS_PPO_CLIP) L = __fmul_rn(step_loss, inv_b);
        else L = __fsub_rn(__fmul_rn(step_loss, inv_b),
                           __fmul_rn(__fmul_rn(__fdiv_rn(1.f, a.hp.focops_lam), __fmul_rn(step_aux0, inv_b)), __fmul_rn(step_aux1, inv_b)));
        acc_loss += static_cast<double>(L);
      }
      step_loss = 0.f; step_aux0 = 0.f; step_aux1 = 0.f;
      }
    }
    PHASE_MARK(8);   // regulariser + sum of squares + block reduction
    // cluster barrier, split: the rows of the next tile are requested while the arrivals propagate
    // (only thread 0 has something to publish -- xchg; a release arrive makes all 512 threads execute a
    //  gpu-scope MEMBAR, 1.1 k cycles per step in the r01 source profile)
    if (tid == 0) asm volatile("fence.acq_rel.cluster;" ::: "memory");
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    stage_next();
    TRACE_MARK(16);
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    PHASE_MARK(9);   // cluster barrier (includes waiting for the slowest net)
    float total = 0.f;
    {
      const unsigned nblk = cluster.num_blocks();
      for (unsigned b = 0; b < nblk; ++b) total += *cluster.map_shared_rank(xchg + par, b);
    }
    // clip coefficient max_norm / (norm + 1e-6), capped at 1 (SFU sqrt and division: <= 2 ulp, and exactly 1 whenever
    // the norm is below the limit)
    const float clip = fminf(__fdividef(a.hp.max_grad_norm, __fadd_rn(sqrt_approx(total), 1e-6f)), 1.f);

    if (active) {
      AdamK k;
      k.w1 = adk[0]; k.b2 = adk[1]; k.w2 = adk[2]; k.ibc2s = adk[3]; k.eps = adk[4]; k.ss = adk[5];
#pragma unroll
      for (int e = 0; e < FE; e += 2) {   // elements e, e+1 are neighbours in a weight row
        int j, kc;
        frag_rc(tid, e, 0, j, kc);
        float2* pw = reinterpret_cast<float2*>(w2 + j * SPO_LDH + kc);
        float2 wv = *pw;
        wv.x = adam_update(wv.x, __fmul_rn(gW2[e], clip), mW2[e], vW2[e], k);
        wv.y = adam_update(wv.y, __fmul_rn(gW2[e + 1], clip), mW2[e + 1], vW2[e + 1], k);
        *pw = wv;
        gW2[e] = 0.f; gW2[e + 1] = 0.f;
        if (kc < D) {   // D and kc even or the pair straddles D: handle the second element on its own
          float* p1 = w1 + j * ldx + kc;
          p1[0] = adam_update(p1[0], __fmul_rn(gW1[0][e], clip), mW1[e], vW1[e], k);
          if (kc + 1 < D) p1[1] = adam_update(p1[1], __fmul_rn(gW1[0][e + 1], clip), mW1[e + 1], vW1[e + 1], k);
        }
        gW1[0][e] = 0.f; gW1[0][e + 1] = 0.f;
        if (NT1 > 1) {
#pragma unroll
          for (int h = 0; h < 2; ++h)
            if (kc + h + 64 < D) {
              float* p1 = w1 + j * ldx + kc + h + 64;
              float m = mv1b[(e + h) * UT + tid], v = mv1b[(FE + e + h) * UT + tid];
              *p1 = adam_update(*p1, __fmul_rn(gW1[NT1 - 1][e + h], clip), m, v, k);
              mv1b[(e + h) * UT + tid] = m;
              mv1b[(FE + e + h) * UT + tid] = v;
            }
          gW1[NT1 - 1][e] = 0.f; gW1[NT1 - 1][e + 1] = 0.f;
        }
      }
      TRACE_MARK(21);
      for (int i = tid; i < SP; i += UT) {
        float* th;
        if (i < SPO_HID) th = b1 + i;
        else if (i < 2 * SPO_HID) th = b2 + (i - SPO_HID);
        else if (i < 2 * SPO_HID + O * SPO_HID) th = w3 + (i - 2 * SPO_HID);
        else if (i < 2 * SPO_HID + O * SPO_HID + O) th = b3 + (i - 2 * SPO_HID - O * SPO_HID);
        else th = log_std + (i - 2 * SPO_HID - O * SPO_HID - O);
        float m = msmall[i], v = vsmall[i];
        *th = adam_update(*th, __fmul_rn(gsmall[i], clip), m, v, k);
        msmall[i] = m; vsmall[i] = v;
        gsmall[i] = 0.f;
      }
    }
    PHASE_MARK(10);  // Adam
    ++step_idx;
    // the __syncthreads at the top of the next iteration orders these weight writes
    // before the next forward
  }
  cp_async_wait_all();
  __syncthreads();

  // ---- write back: weights, moments, step counters, logged losses ----
  if (active) {
    for (int i = tid; i < SPO_HID * D; i += UT) {
      const int j = i / D, kx = i - j * D;
      a.params[off.w1 + i] = w1[j * ldx + kx];
    }
    for (int i = tid; i < SPO_HID * SPO_HID; i += UT)
      a.params[off.w2 + i] = w2[(i >> 6) * SPO_LDH + (i & 63)];
    for (int i = tid; i < SP; i += UT) {
      float th;
      if (i < SPO_HID) th = b1[i];
      else if (i < 2 * SPO_HID) th = b2[i - SPO_HID];
      else if (i < 2 * SPO_HID + O * SPO_HID) th = w3[i - 2 * SPO_HID];
      else if (i < 2 * SPO_HID + O * SPO_HID + O) th = b3[i - 2 * SPO_HID - O * SPO_HID];
      else th = log_std[i - 2 * SPO_HID - O * SPO_HID - O];
      const int g = sm.goff(off, i);
      a.params[g] = th;
      a.adam_m[g] = msmall[i];
      a.adam_v[g] = vsmall[i];
    }
#pragma unroll
    for (int e = 0; e < FE; ++e) {
      int j, kc;
      frag_rc(tid, e, 0, j, kc);
      a.adam_m[off.w2 + j * SPO_HID + kc] = mW2[e];
      a.adam_v[off.w2 + j * SPO_HID + kc] = vW2[e];
      if (kc < D) {
        a.adam_m[off.w1 + j * D + kc] = mW1[e];
        a.adam_v[off.w1 + j * D + kc] = vW1[e];
      }
      if (NT1 > 1 && kc + 64 < D) {
        a.adam_m[off.w1 + j * D + kc + 64] = mv1b[e * UT + tid];
        a.adam_v[off.w1 + j * D + kc + 64] = mv1b[(FE + e) * UT + tid];
      }
    }
    if (tid == 0) {
      a.adam_t[net] = t0 + static_cast<int>(n_steps);
      const int slot = (net == 0) ? 2 : (net == 1 ? 0 : 1);
      atomicAdd(&ctrl->loss_sum[slot], acc_loss);
    }
  }
#ifdef SPO_PHASE_TIMERS_V1
  if (tid < 16) atomicAdd(&g_phase_cycles[rank & 3][tid], sm_phase__[tid]);
#endif
  if (rank == 1 && tid == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&ctrl->steps), static_cast<unsigned long long>(n_steps));
  cluster.sync();  // no CTA may exit while a peer can still read its shared memory
}

size_t update_smem_bytes(int nt1) {
  const int ldx = upd_ldx(nt1);
  size_t f = 4 * SPO_ROWS + 4 * SPO_MAX_ACT + 8 + SPO_HID * ldx + SPO_HID + SPO_HID * SPO_LDH + SPO_HID + SPO_MAX_ACT * SPO_HID + SPO_MAX_ACT + 8 +
             3 * SPN + SPO_ROWS * ldx + SPO_ROWS * AUXW + 3 * SPO_ROWS * SPO_LDH + 3 * SPO_ROWS * SPO_MAX_ACT + 64 + 4 + (nt1 > 1 ? 2 * FE * UT : 0);
  return f * sizeof(float);
}

template <int NT1>
int launch_update(const UpdArgsV1& a, cudaStream_t stream) {
  const size_t smem = update_smem_bytes(NT1);
  SPO_REQUIRE(smem <= 227 * 1024, SPO_ERR_UNSUPPORTED, "spo_pg_update: obs_dim=%d needs %zu B of shared memory (> 227 KB)", a.D, smem);
  static int cluster_size = 0;   // 4 preferred (barrier measured faster than for 3, profiles/r01_ubench.txt); 3 as fallback
  SPO_CUDA_TRY(cudaFuncSetAttribute(spo_update_v1_kernel<NT1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  if (!cluster_size) {
    const char* env = getenv("SPO_CLUSTER");   // debugging aid: pin the cluster size (3 or 4)
    if (env && (env[0] == '3' || env[0] == '4')) cluster_size = env[0] - '0';
  }
  for (int attempt = 0; attempt < 2; ++attempt) {
    const int cs = cluster_size ? cluster_size : (attempt == 0 ? 4 : 3);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cs);
    cfg.blockDim = dim3(UT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, spo_update_v1_kernel<NT1>, a);
    if (e == cudaSuccess) { cluster_size = cs; return SPO_OK; }
    if (cluster_size || attempt == 1) {
      spo_set_error("spo_pg_update: launch failed (cluster=%d): %s", cs, cudaGetErrorString(e));
      return SPO_ERR_CUDA;
    }
    (void)cudaGetLastError();  // clear and retry with the other cluster size
  }
  return SPO_ERR_CUDA;
}

}  // namespace

#ifdef SPO_PHASE_TIMERS_V1
extern "C" int spo_debug_phase_cycles(unsigned long long* out64, int reset) {
  SPO_CUDA_TRY(cudaMemcpyFromSymbol(out64, g_phase_cycles, sizeof(unsigned long long) * 64));
  if (reset) {
    unsigned long long z[64] = {0};
    SPO_CUDA_TRY(cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z)));
  }
  return SPO_OK;
}
extern "C" int spo_debug_trace(long long* out_4x16x24) {
  SPO_CUDA_TRY(cudaMemcpyFromSymbol(out_4x16x24, g_trace, sizeof(long long) * 4 * 16 * 24));
  return SPO_OK;
}
#endif

extern "C" int spo_comm_slot_floats_v1(const spo_dims* d, int* slot_floats) {
  int rc = spo_check_dims(d);
  if (rc) return rc;
  SPO_REQUIRE(slot_floats, SPO_ERR_INVALID_ARG, "spo_comm_slot_floats: null output");
  const int nt1 = d->obs_dim <= 64 ? 1 : 2;
  *slot_floats = UT * FE * (1 + nt1) + spo_pad4(2 * SPO_HID + d->act_dim * SPO_HID + 2 * d->act_dim);
  return SPO_OK;
}

extern "C" int spo_pg_update_v1_dp(const spo_dims* d, float* params, float* adam_m, float* adam_v, int* adam_t,
                                   const spo_batch* data, const int64_t* perm, int64_t perm_len, int batch,
                                   spo_loss_kind kind, const spo_hparams* hp, spo_update_ctrl* ctrl,
                                   const spo_comm* comm, void* stream);

extern "C" int spo_pg_update_v1(const spo_dims* d, float* params, float* adam_m, float* adam_v, int* adam_t,
                             const spo_batch* data, const int64_t* perm, int64_t perm_len, int batch,
                             spo_loss_kind kind, const spo_hparams* hp, spo_update_ctrl* ctrl, void* stream) {
  return spo_pg_update_v1_dp(d, params, adam_m, adam_v, adam_t, data, perm, perm_len, batch, kind, hp, ctrl, nullptr, stream);
}

extern "C" int spo_pg_update_v1_dp(const spo_dims* d, float* params, float* adam_m, float* adam_v, int* adam_t,
                                const spo_batch* data, const int64_t* perm, int64_t perm_len, int batch,
                                spo_loss_kind kind, const spo_hparams* hp, spo_update_ctrl* ctrl,
                                const spo_comm* comm, void* stream) {
  int rc = spo_check_dims(d);
  if (rc) return rc;
  SPO_REQUIRE(params && adam_m && adam_v && adam_t && data && perm && hp && ctrl, SPO_ERR_INVALID_ARG, "spo_pg_update: null argument");
  SPO_REQUIRE(batch > 0 && perm_len > 0 && perm_len <= data->count, SPO_ERR_INVALID_ARG,
              "spo_pg_update: batch=%d perm_len=%lld count=%lld", batch, (long long)perm_len, (long long)data->count);
  SPO_REQUIRE(kind >= SPO_LOSS_PPO_CLIP && kind <= SPO_LOSS_CUP_PROJECTION, SPO_ERR_INVALID_ARG, "spo_pg_update: kind=%d", (int)kind);
  SPO_REQUIRE(data->obs && data->target_r && data->target_c, SPO_ERR_INVALID_ARG, "spo_pg_update: batch obs/targets null");
  if (kind != SPO_LOSS_CRITIC_ONLY)
    SPO_REQUIRE(data->act && data->logp && data->adv, SPO_ERR_INVALID_ARG, "spo_pg_update: actor loss needs act/logp/adv");
  if (kind == SPO_LOSS_FOCOPS || kind == SPO_LOSS_CUP_PROJECTION) {
    SPO_REQUIRE(data->old_mean && data->old_std, SPO_ERR_INVALID_ARG, "spo_pg_update: FOCOPS needs old_mean/old_std");
    SPO_REQUIRE(batch <= SPO_ROWS, SPO_ERR_UNSUPPORTED, "spo_pg_update: FOCOPS supports batch <= %d (got %d)", SPO_ROWS, batch);
  }
  UpdArgsV1 a{};
  a.params = params; a.adam_m = adam_m; a.adam_v = adam_v; a.adam_t = adam_t;
  a.data = *data; a.perm = perm; a.perm_len = perm_len; a.batch = batch; a.kind = kind;
  a.D = d->obs_dim; a.A = d->act_dim; a.hp = *hp; a.ctrl = ctrl;
  if (kind == SPO_LOSS_CUP_PROJECTION) {
    // (c * ratio * adv [B] + kl [B,1]).mean() = mean(kl) + c * mean(ratio * adv): the FOCOPS loss
    // (kl - ratio * adv / lam) * 1(kl <= delta) with delta = inf (mask 1) and 1/lam = -c
    a.kind = SPO_LOSS_FOCOPS;
    a.actor_only = 1;
    a.hp.focops_kl = INFINITY;
    a.hp.focops_lam = -1.f / hp->focops_lam;   // c = 0 -> -inf -> 1/lam = -0
  }
  if (kind == SPO_LOSS_PG) {
    // the unclipped surrogate is the clipped one with an unbounded clip range: clamp(ratio) == ratio, the
    // min() keeps the first branch, value and gradient are those of pg.py:309 bit for bit
    a.kind = SPO_LOSS_PPO_CLIP;
    a.hp.clip_lo = -INFINITY;
    a.hp.clip_hi = INFINITY;
  }
  if (comm && comm->world > 1) {
    SPO_REQUIRE(comm->rank >= 0 && comm->rank < comm->world && comm->world <= 32 && comm->grad_bufs && comm->flags,
                SPO_ERR_INVALID_ARG, "spo_pg_update_dp: bad spo_comm (world=%d rank=%d)", comm->world, comm->rank);
    a.comm = *comm;
  } else {
    a.comm.world = 1;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d->obs_dim <= 64) return launch_update<1>(a, st);
  return launch_update<2>(a, st);
}
