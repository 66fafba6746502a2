// w2a_kernels.cuh -- sm_100a kernels of the HeatAlertEnv hot path.
//
// What the kernels compute is the reference's HeatAlertEnv.step / reset
// (src/weather2alert/env.py:133-262 of NSAPH-Projects/weather2alert) for a whole batch
// of independent envs; how they compute it is laid out for B200:
//
//   * per-env state is two SoA planes of 8 bytes (static: episode row + coefficient
//     vector; mutable: remaining budget + packed day/streak/14-day window), so a warp
//     reads and writes it with fully coalesced 8-byte accesses;
//   * the step kernel gives every env 8 lanes: the env's three table rows (float64
//     reward features, float32 observation slots, float32 coefficient vector) are each
//     one contiguous 128-256 B run read with 128-bit loads by those 8 lanes, the two
//     28-term dot products are accumulated in float64 and folded with warp shuffles,
//     and the 29-slot observation goes out as four contiguous 32-byte pieces;
//   * tables are read through the read-only path and stay L2/L1 resident; the per-env
//     streams (state, action, obs, reward, flags) use streaming (.cs) accesses so they
//     do not evict the tables.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace w2a {

constexpr int kObsDim = 29;
constexpr int kXrewStride = 24;
constexpr int kXobsStride = 32;
constexpr int kCoefStride = 64;
constexpr unsigned kFull = 0xffffffffu;
// env.py:221 -- the reference divides by 152, not 153
constexpr double kRewardScale = -1000.0 / 152.0;

struct Tables {
  const double* xrew;      // [R, 24] float64
  const float* xobs;       // [R, 32] float32, float 4*q+i = slot 8*i+q
  const uint8_t* n_days;   // [S*Y]
  const int32_t* budget0;  // [S*Y]
  const float* w;          // [Sw*P, 64] float32, lane-permuted
  int S, Y, T, P, Sw;
};

struct StepArgs {
  Tables tb;
  const uint2* st_static;  // x: row0 (24 bit) | n_days << 24 ; y: coefficient vector index
  uint2* st_mut;           // x: remaining budget (int32)      ; y: t | streak<<8 | win14<<16 | done<<30
  const void* action;
  float* obs;
  float* reward;
  uint8_t* terminated;
  int32_t* remaining;
  uint8_t* at_budget;
  int n;
};

struct ResetArgs {
  Tables tb;
  const int32_t* loc;
  const int32_t* year;
  const int32_t* coef_col;
  const int32_t* coef_idx;
  const int32_t* budget;  // may be null
  const uint8_t* mask;    // may be null
  uint2* st_static;
  uint2* st_mut;
  float* obs;
  int32_t* remaining;  // may be null
  int n;
};

// ---- memory-access helpers ------------------------------------------------------------
__device__ __forceinline__ uint2 ld_stream(const uint2* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(uint2* p, uint2 v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }

template <int ADT>
__device__ __forceinline__ int load_action(const void* a, int e) {
  if (ADT == 0) return __ldcs(reinterpret_cast<const unsigned char*>(a) + e) != 0;
  if (ADT == 1) return __ldcs(reinterpret_cast<const int*>(a) + e) != 0;
  return __ldcs(reinterpret_cast<const long long*>(a) + e) != 0;
}

// scipy.special.expit in float64, as used at env.py:211,218
__device__ __forceinline__ double expit(double z) { return 1.0 / (1.0 + exp(-z)); }

// position of a coefficient inside the 64-float lane-permuted vector
//   table feature u (0..23: slots 0..22 and 26), head h: u < 16 -> 2u+h ; else 32+4(u-16)+h
//   lane-q "other" term (q=0 lag1, 1 streak, 2 remaining budget, 3 bias), head h: 32+4q+2+h
__host__ __device__ __forceinline__ int coef_pos_table(int u, int h) {
  return u < 16 ? 2 * u + h : 32 + 4 * (u - 16) + h;
}
__host__ __device__ __forceinline__ int coef_pos_other(int q, int h) { return 32 + 4 * q + 2 + h; }
__host__ __device__ __forceinline__ int xobs_pos(int slot) { return 4 * (slot & 7) + (slot >> 3); }

// ---- the integer alert-history transition (env.py:238-250, 256-260) ---------------------
struct Transition {
  int act;        // clamped action actually applied (env.py:243-246)
  int at_budget;  // pre-action flag (env.py:242; info["at_budget"])
  int rem;        // remaining budget after the action (env.py:193,249-250)
  int lag;        // alert_lag1 seen by reward and obs: today's action unless t == 0 (env.py:190; Q1)
  int streak;     // alert_streak seen by reward and obs: yesterday's streak (env.py:192; Q1)
  int cnt14;      // alert_2wks: agent alerts in the last 14 steps, today included (env.py:191)
  int done;       // t >= n_days - 1 (env.py:256)
  unsigned row;   // table row of day t
};

__device__ __forceinline__ Transition advance(uint2 sv, uint2& mv, int a) {
  Transition o;
  int rem = static_cast<int>(mv.x);
  const unsigned pk = mv.y;
  const int t = pk & 0xff;
  const int streak = (pk >> 8) & 0xff;
  unsigned win = (pk >> 16) & 0x3fffu;
  const int n_days = sv.x >> 24;
  o.at_budget = (rem == 0);  // sum(actual) == budget  <=>  budget - sum(actual) == 0
  o.act = (a && o.at_budget) ? 0 : a;
  rem -= o.act;
  win = ((win << 1) | static_cast<unsigned>(o.act)) & 0x3fffu;
  o.rem = rem;
  o.lag = t > 0 ? o.act : 0;
  o.streak = streak;
  o.cnt14 = __popc(win);
  o.done = (t >= n_days - 1);
  o.row = (sv.x & 0xffffffu) + static_cast<unsigned>(t);
  // Stepping past the end is legal (Q11): the buffers keep growing (rem, win) while
  // t and the streak freeze (env.py:257-260).
  const int t2 = o.done ? t : t + 1;
  const int s2 = o.done ? streak : (o.act ? streak + 1 : 0);
  mv.x = static_cast<unsigned>(rem);
  mv.y = static_cast<unsigned>(t2) | (static_cast<unsigned>(s2) << 8) | (win << 16) |
         (static_cast<unsigned>(o.done) << 30);
  return o;
}

// reward = -1000/152 * sigmoid(z_b) * (1 - sigmoid(z_e) * [heat_qi > 0.5] * action)   (env.py:211-221)
__device__ __forceinline__ float reward_from_logits(double zb, double ze, int gate, int act) {
  const double baseline = expit(zb);
  const double eff = (gate && act) ? expit(ze) : 0.0;
  return static_cast<float>(kRewardScale * baseline * (1.0 - eff));
}

// =========================================================================================
// step, variant THREAD: one thread per env.  The straightforward kernel: used for small
// batches (launch-latency bound anyway) and as the in-tree cross-check of the COOP kernel.
// =========================================================================================
template <int ADT>
__global__ void __launch_bounds__(128) step_thread_kernel(const StepArgs p) {
  const int stride = gridDim.x * blockDim.x;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < p.n; e += stride) {
    const uint2 sv = ld_stream(p.st_static + e);
    uint2 mv = ld_stream(p.st_mut + e);
    const int a = load_action<ADT>(p.action, e);
    const Transition tr = advance(sv, mv, a);
    st_stream(p.st_mut + e, mv);

    const double* __restrict__ xr = p.tb.xrew + static_cast<size_t>(tr.row) * kXrewStride;
    const float* __restrict__ wv = p.tb.w + static_cast<size_t>(sv.y) * kCoefStride;
    double zb = 0.0, ze = 0.0;
#pragma unroll
    for (int u = 0; u < 24; ++u) {
      const double x = __ldg(xr + u);
      zb = fma(x, static_cast<double>(__ldg(wv + coef_pos_table(u, 0))), zb);
      ze = fma(x, static_cast<double>(__ldg(wv + coef_pos_table(u, 1))), ze);
    }
    const double other[4] = {static_cast<double>(tr.lag), static_cast<double>(tr.streak),
                             static_cast<double>(tr.rem), 1.0};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      zb = fma(other[q], static_cast<double>(__ldg(wv + coef_pos_other(q, 0))), zb);
      ze = fma(other[q], static_cast<double>(__ldg(wv + coef_pos_other(q, 1))), ze);
    }
    const int gate = __ldg(xr) > 0.5;  // heat_qi is table feature 0 (env.py:218; Q4)
    p.reward[e] = reward_from_logits(zb, ze, gate, tr.act);
    p.terminated[e] = static_cast<uint8_t>(tr.done);
    if (p.remaining) p.remaining[e] = tr.rem;
    if (p.at_budget) p.at_budget[e] = static_cast<uint8_t>(tr.at_budget);

    if (!tr.done) {  // on and after the last day the previous observation stands (Q2)
      const float* __restrict__ xo = p.tb.xobs + static_cast<size_t>(tr.row) * kXobsStride;
      float* ob = p.obs + static_cast<size_t>(e) * kObsDim;
#pragma unroll
      for (int s = 0; s < kObsDim; ++s) {
        float v;
        if (s == 23) v = static_cast<float>(tr.lag);
        else if (s == 24) v = static_cast<float>(tr.streak);
        else if (s == 25) v = static_cast<float>(tr.rem);
        else if (s == 28) v = static_cast<float>(tr.cnt14);
        else v = __ldg(xo + xobs_pos(s));
        ob[s] = v;
      }
    }
  }
}

// =========================================================================================
// step, variant COOP: a warp owns 32 consecutive envs.  Phase A is one thread per env
// (coalesced state/action traffic, integer transition).  Phase B walks the 32 envs four
// at a time, 8 lanes per env: every table row is fetched as one contiguous run, the
// float64 partial dot products are folded inside the 8-lane group with shuffles, and the
// observation row is written as contiguous 32-byte pieces.  Phase C is one thread per
// env again (two float64 sigmoids, reward, flags).
// =========================================================================================
template <int ADT>
__global__ void __launch_bounds__(256) step_coop_kernel(const StepArgs p) {
  const int lane = threadIdx.x & 31;
  const int q = lane & 7;
  const bool hi = (q & 4) != 0;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const int n_groups = (p.n + 31) >> 5;

  for (int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; wg < n_groups; wg += warps_total) {
    const int base = wg << 5;
    const int e = base + lane;
    const bool valid = e < p.n;
    const int ec = valid ? e : p.n - 1;

    // ---- phase A: state transition, one thread per env ----------------------------------
    const uint2 sv = ld_stream(p.st_static + ec);
    uint2 mv = ld_stream(p.st_mut + ec);
    const int a = load_action<ADT>(p.action, ec);
    const Transition tr = advance(sv, mv, a);
    if (valid) st_stream(p.st_mut + e, mv);
    // lag:1 | skip_obs:1 | streak:8 | cnt14:8 -- what the other lanes need of this env
    const int packed = tr.lag | ((tr.done || !valid) ? 2 : 0) | (tr.streak << 8) | (tr.cnt14 << 16);

    // ---- phase B: 8 lanes per env ---------------------------------------------------------
    double zb = 0.0, ze = 0.0;
    int gate = 0;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const unsigned rowj = __shfl_sync(kFull, tr.row, it, 8);
      const unsigned cvj = __shfl_sync(kFull, sv.y, it, 8);
      const int pkj = __shfl_sync(kFull, packed, it, 8);
      const int remj = __shfl_sync(kFull, tr.rem, it, 8);

      const double* __restrict__ xr = p.tb.xrew + static_cast<size_t>(rowj) * kXrewStride;
      const double2 x01 = __ldg(reinterpret_cast<const double2*>(xr) + q);  // features 2q, 2q+1
      const double x2 = __ldg(xr + 16 + q);                                 // feature 16+q
      const float4* __restrict__ wv =
          reinterpret_cast<const float4*>(p.tb.w + static_cast<size_t>(cvj) * kCoefStride);
      const float4 w0 = __ldg(wv + q);      // (b,e) of feature 2q, (b,e) of feature 2q+1
      const float4 w1 = __ldg(wv + 8 + q);  // (b,e) of feature 16+q, (b,e) of this lane's other term
      const float4 xo =
          __ldg(reinterpret_cast<const float4*>(p.tb.xobs + static_cast<size_t>(rowj) * kXobsStride) + q);

      const int lagj = pkj & 1;
      const int streakj = (pkj >> 8) & 0xff;
      // the term that does not come from the table: lag1 / streak / remaining / bias / none
      const double xoth = q == 0 ? static_cast<double>(lagj)
                        : q == 1 ? static_cast<double>(streakj)
                        : q == 2 ? static_cast<double>(remj)
                        : q == 3 ? 1.0 : 0.0;
      double pb = x01.x * static_cast<double>(w0.x);
      double pe = x01.x * static_cast<double>(w0.y);
      pb = fma(x01.y, static_cast<double>(w0.z), pb);
      pe = fma(x01.y, static_cast<double>(w0.w), pe);
      pb = fma(x2, static_cast<double>(w1.x), pb);
      pe = fma(x2, static_cast<double>(w1.y), pe);
      pb = fma(xoth, static_cast<double>(w1.z), pb);
      pe = fma(xoth, static_cast<double>(w1.w), pe);
      const int g0 = (q == 0) && (x01.x > 0.5);  // heat_qi gate lives in lane 0 (Q4)

      // fold 8 lanes x 2 heads: lanes 0-3 end with z_b, lanes 4-7 with z_e
      double v = (hi ? pe : pb) + __shfl_xor_sync(kFull, hi ? pb : pe, 4);
      v += __shfl_xor_sync(kFull, v, 2);
      v += __shfl_xor_sync(kFull, v, 1);
      const double o = __shfl_xor_sync(kFull, v, 4);
      const int gj = __shfl_sync(kFull, g0, 0, 8);
      if (q == it) {
        zb = hi ? o : v;
        ze = hi ? v : o;
        gate = gj;
      }

      if (!(pkj & 2)) {  // observation of day t (Q2): table row + the four agent slots
        float v2 = xo.z, v3 = xo.w;
        if (q == 7) v2 = static_cast<float>(lagj);                  // slot 23 alert_lag1
        if (q == 0) v3 = static_cast<float>(streakj);               // slot 24 alert_streak
        if (q == 1) v3 = static_cast<float>(remj);                  // slot 25 remaining_budget
        if (q == 4) v3 = static_cast<float>((pkj >> 16) & 0xff);    // slot 28 alert_2wks
        float* ob = p.obs + static_cast<size_t>(base + (lane & 24) + it) * kObsDim + q;
        st_stream(ob, xo.x);
        st_stream(ob + 8, xo.y);
        st_stream(ob + 16, v2);
        if (q <= 4) st_stream(ob + 24, v3);
      }
    }

    // ---- phase C: reward and flags, one thread per env --------------------------------------
    if (valid) {
      __stcs(p.reward + e, reward_from_logits(zb, ze, gate, tr.act));
      p.terminated[e] = static_cast<uint8_t>(tr.done);
      if (p.remaining) __stcs(p.remaining + e, tr.rem);
      if (p.at_budget) p.at_budget[e] = static_cast<uint8_t>(tr.at_budget);
    }
  }
}

// =========================================================================================
// reset: host index streams -> state + first observation (env.py:155-184).
// Same warp shape as the COOP step: thread per env for the state, 8 lanes per env for
// the observation row.
// =========================================================================================
__global__ void __launch_bounds__(256) reset_kernel(const ResetArgs p) {
  const int lane = threadIdx.x & 31;
  const int q = lane & 7;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const int n_groups = (p.n + 31) >> 5;
  for (int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; wg < n_groups; wg += warps_total) {
    const int base = wg << 5;
    const int e = base + lane;
    const bool valid = e < p.n;
    const int ec = valid ? e : p.n - 1;
    const bool take = valid && (p.mask == nullptr || p.mask[ec] != 0);

    const int sy = __ldcs(p.loc + ec) * p.tb.Y + __ldcs(p.year + ec);
    const unsigned nd = __ldg(p.tb.n_days + sy);
    const unsigned row0 = static_cast<unsigned>(sy) * static_cast<unsigned>(p.tb.T);
    int budget = p.budget ? __ldcs(p.budget + ec) : INT32_MIN;
    if (budget == INT32_MIN) budget = __ldg(p.tb.budget0 + sy);  // env.py:167-170
    const unsigned cv = static_cast<unsigned>(__ldcs(p.coef_col + ec)) * static_cast<unsigned>(p.tb.P) +
                        static_cast<unsigned>(__ldcs(p.coef_idx + ec));
    if (take) {
      st_stream(p.st_static + e, make_uint2(row0 | (nd << 24), cv));
      st_stream(p.st_mut + e, make_uint2(static_cast<unsigned>(budget), 0u));  // t = streak = win = 0
      if (p.remaining) p.remaining[e] = budget;
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const unsigned rowj = __shfl_sync(kFull, row0, it, 8);
      const int budj = __shfl_sync(kFull, budget, it, 8);
      const int takej = __shfl_sync(kFull, static_cast<int>(take), it, 8);
      if (takej) {
        const float4 xo =
            __ldg(reinterpret_cast<const float4*>(p.tb.xobs + static_cast<size_t>(rowj) * kXobsStride) + q);
        float v2 = xo.z, v3 = xo.w;
        if (q == 7) v2 = 0.f;                          // alert_lag1 (t == 0)
        if (q == 0) v3 = 0.f;                          // alert_streak
        if (q == 1) v3 = static_cast<float>(budj);     // remaining_budget = budget
        if (q == 4) v3 = 0.f;                          // alert_2wks
        float* ob = p.obs + static_cast<size_t>(base + (lane & 24) + it) * kObsDim + q;
        st_stream(ob, xo.x);
        st_stream(ob + 8, xo.y);
        st_stream(ob + 16, v2);
        if (q <= 4) st_stream(ob + 24, v3);
      }
    }
  }
}

}  // namespace w2a
