// w2a_kernels.cuh -- sm_100a kernels of the HeatAlertEnv hot path.
//
// What the kernels compute is the reference's HeatAlertEnv.step / reset
// (src/weather2alert/env.py:133-262 of NSAPH-Projects/weather2alert) for a whole batch
// of independent envs; how they compute it is laid out for B200:
//
//   * per-env state is two SoA planes of 8 bytes (static: episode row + coefficient
//     record + year; mutable: remaining budget + packed day/streak/14-day window), so a warp
//     reads and writes it with fully coalesced 8-byte accesses;
//   * everything in the reward that does not depend on the agent is tabulated once for the
//     whole summer (w2a_precompute_logits: a float64 GEMM per county and day on the FP64
//     tensor path), so the default step kernel (TABLE) gathers one 32-byte sector and one
//     128-byte float32 observation row per env, and the episode-fused rollout streams 16
//     bytes per env-step;
//   * the cooperative step kernel (COOP: env sets on similar counties' coefficients, which
//     cannot be tabulated) gathers the two 256-byte, line-aligned records of an env -- one
//     day-row of float64 features, one float32 coefficient vector: exactly four 128-byte
//     lines, each fetched by 8 lanes with one 128-bit load apiece -- and accumulates the two
//     28-term dot products in float64, folding the 8 partials through shared memory;
//   * the 29-slot observations of a warp's 32 envs are staged in shared memory so that they
//     leave as one contiguous 3712-byte bulk (TMA) store;
//   * tables are read through the read-only path; the per-env streams (action, obs, reward,
//     flags) use streaming accesses, the 16 bytes of state per env stay L2-resident between
//     two steps (evict_last) when the env set is small enough.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace w2a {

constexpr int kObsDim = 29;
constexpr int kXStride = 32;   // float64 per day-row record
constexpr int kWStride = 64;   // float32 per coefficient record
constexpr int kXSig = 24;      // record position of the significance code (0 = none)
constexpr int kWOther = 56;    // first float of the non-table coefficients
constexpr unsigned kFull = 0xffffffffu;
constexpr unsigned kRecMask = 0xffffffu;  // static state plane, word y: coefficient record index; the year index sits above it
// env.py:221 -- the reference divides by 152, not 153
constexpr double kRewardScale = -1000.0 / 152.0;

struct Tables {
  // [R, 32] float64: 0..23 = reward features read from the table (obs slots 0..22 and 26),
  // 24 = significance code (0.0 = none, k = label k), 25 = heat_qi gate (0.0 if heat_qi > 0.5
  // else -1e300), 26..31 = 0.0
  const double* x;
  const uint8_t* n_days;   // [S*Y]
  const int32_t* budget0;  // [S*Y]
  // [Sw*P, 64] float32: (baseline, effectiveness) pairs; 0..47 = the 24 table features,
  // 48,49 = 0, 50,51 = (0, 1) the gate column, 52..55 = 0, 56..63 = alert_lag1, alert_streak,
  // remaining_budget, bias
  const float* w;
  int S, Y, T, P, Sw;
  const int32_t* col_to_loc;  // [Sw] feature-store county of coefficient column `col`, -1 = none (w2a_set_county_map)
  // ---- whole-summer tables (w2a_precompute_logits), null / 0 before ------------------------------------
  // For every day t, coefficient record (col*P + p) and year of the column's county: the agent-independent part
  // of both logits (24 table features, bias and heat_qi gate; env.py:206-217), in two forms:
  // z32_all[t][col*P + p][year]: ONE 32-byte sector = the six agent coefficients (lag, streak, remaining budget as
  // (b, e) pairs) + (z_b, z_e) rounded to float32 -- everything an env needs for its logits in one L2 request of
  // the step kernel.  The float32 rounding (<= 1e-6 below |z| = 32) is inside the reward tolerance; where it could
  // matter (|z| >= 32 that the agent terms pull back into the sigmoid's sensitive range) the kernel takes the
  // float64 pair instead.
  const float4* z32_all;
  // zt_all[(col*P + p)*Y + year][t]: (z_b, z_e) float64, episode-major: the T days of one (record, year) are one
  // contiguous run -- what the episode-fused rollout streams (16 B per env-step instead of the 256-byte day-row and
  // a 24-term dot product), and the step kernel's float64 fallback.
  const double2* zt_all;
  unsigned n_records;         // Sw * P
  // pt[row]: the closed-loop policy's table logit of day-row `row` (w2a_rollout_policy computes it per call)
  double* pt;
  // x32[row][32]: the float32 observation row (slot order 0..28, significance as NaN / label) of
  // day-row `row` of x: the row copy of _get_obs (env.py:187) is one 128-byte line per env
  const float* x32;
};

struct StepArgs {
  Tables tb;
  const uint2* st_static;  // x: row0 (24 bit) | n_days << 24 ; y: coefficient record index (24 bit) | year index << 24
  uint2* st_mut;           // x: remaining budget (int32)      ; y: t | streak<<8 | win14<<16 | done<<30 | canonical<<31
  const void* action;
  float* obs;
  float* reward;
  uint8_t* terminated;
  int32_t* remaining;
  uint8_t* at_budget;
  int n;
  int keep_state;  // the two state planes (16 B/env) are small enough to live in L2 from one step to the next
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

// ---- helpers --------------------------------------------------------------------------
// programmatic dependent launch: a step kernel lets the next launch on the stream get resident while it
// drains, and itself waits for the previous launch (and its memory) before touching anything it wrote
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ uint2 ld_stream(const uint2* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(uint2* p, uint2 v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }

template <int ADT>
__device__ __forceinline__ int load_action(const void* a, int e) {
  if (ADT == 0) return __ldcs(reinterpret_cast<const unsigned char*>(a) + e) != 0;
  if (ADT == 1) return __ldcs(reinterpret_cast<const int*>(a) + e) != 0;
  return __ldcs(reinterpret_cast<const long long*>(a) + e) != 0;
}

// scipy.special.expit in float64, as used at env.py:211,218 (the episode-fused rollout sums float64 returns)
__device__ __forceinline__ double expit(double z) { return 1.0 / (1.0 + exp(-z)); }

// The same function for the episode-fused kernels, where two (policy: up to ten) of them per env-step are the bulk of
// the arithmetic: exp(-z) = 2^n * e^g with n = rint(-z log2 e), g = (-z log2 e - n) ln 2 in [-0.347, 0.347] and a
// degree-10 Taylor polynomial (truncation 2e-13 relative), 2^n by exponent arithmetic, 1 / (1 + e) by the hardware
// reciprocal seed and two Newton steps: ~25 instructions instead of ~60 for exp() + a float64 division.  Relative
// error <= 4e-13 for every z (returns are checked against the oracle at 1e-9); z = -1e300 (the closed heat_qi gate)
// gives 2^-1000 (1 - that is exactly 1), z = +1e300 exactly 1, NaN stays NaN.
__device__ __forceinline__ double expit_fast(double z) {
  double t = z * -1.4426950408889634;                 // 2^t = exp(-z)
  t = fmin(fmax(t, -1000.0), 1000.0);                 // 2^+-1000 are finite doubles; the sigmoid is 0 / 1 out there
  const double magic = 6755399441055744.0;            // 1.5 * 2^52: t + magic holds rint(t) in its low word
  const double tn = t + magic;
  const int n = __double2loint(tn);
  const double g = (t - (tn - magic)) * 0.6931471805599453;
  double p = 2.755731922398589e-07;                   // 1/10!
  p = fma(p, g, 2.7557319223985893e-06);              // 1/9!
  p = fma(p, g, 2.48015873015873e-05);                // 1/8!
  p = fma(p, g, 1.984126984126984e-04);               // 1/7!
  p = fma(p, g, 1.388888888888889e-03);               // 1/6!
  p = fma(p, g, 8.333333333333333e-03);               // 1/5!
  p = fma(p, g, 4.1666666666666664e-02);              // 1/4!
  p = fma(p, g, 1.6666666666666666e-01);              // 1/3!
  p = fma(p, g, 0.5);
  p = fma(p, g, 1.0);
  p = fma(p, g, 1.0);                                 // e^g
  const double e = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));  // * 2^n (p in [0.7, 1.42])
  const double d = 1.0 + e;
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  r = fma(r, fma(-d, r, 1.0), r);                     // Newton: r (2 - d r), twice: 20 -> 40 -> 80 bits
  r = fma(r, fma(-d, r, 1.0), r);
  return r;
}

// The step kernels hand the reward back as float32, so they evaluate the sigmoids in float32 too --
// but on an argument reduced in float64: t = -z*log2(e) = n + f exactly (|f| <= 1/2), 2^f by
// MUFU.EX2, 2^n by exponent arithmetic, the reciprocal by MUFU.RCP.  Relative error <= 3e-7 for any
// z (the reference tolerance is 1e-5); ~14 instructions instead of ~60 for the float64 exp + divide.
__device__ __forceinline__ float sigmoid_f32(double z) {
  double t = z * -1.4426950408889634;          // 2^t = exp(-z)
  t = fmin(fmax(t, -126.0), 126.0);            // keeps 2^t a finite float32; sigmoid is 1 / 1e-38 out there
  const double magic = 6755399441055744.0;     // 1.5 * 2^52: t + magic holds rint(t) in its low word
  const double tn = t + magic;
  const int n = __double2loint(tn);
  const float f = static_cast<float>(t - (tn - magic));
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(f));
  const float p = e * __int_as_float((n + 127) << 23);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + p));
  return r;
}
// reward = -1000/152 * sigmoid(z_b) * (1 - sigmoid(z_e) * [heat_qi > 0.5] * action)  (env.py:211-221).
// 1 - sigmoid(z_e) is taken as sigmoid(-z_e), which has no cancellation; the reference's float64
// `1 - expit(z_e)` does cancel, visibly (beyond 1e-5 relative) once z_e > ~25, so from z_e > 20 on the
// reference's own float64 expression is evaluated instead (a branch real coefficients never take) and
// the result follows the reference's rounding there as well.  The gate column makes z_e = -1e300 when
// heat_qi <= 0.5, and sigmoid(1e300) is exactly 1.
__device__ __forceinline__ float step_reward(double zb, double ze, int act) {
  const float b = sigmoid_f32(zb);
  float q = 1.0f;
  if (act) q = ze > 20.0 ? static_cast<float>(1.0 - expit(ze)) : sigmoid_f32(-ze);
  return static_cast<float>(kRewardScale) * b * q;
}

// observation slot of table feature u (u = 0..23): slots 0..22, then 26 (issued_in_advance)
__host__ __device__ __forceinline__ int slot_of_feature(int u) { return u < 23 ? u : 26; }
// significance: record value 0.0 means "no alert" and is shown as NaN (the reference's
// obs carries NaN there)
__device__ __forceinline__ float sig_to_obs(float code) { return code == 0.f ? CUDART_NAN_F : code; }

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
         (static_cast<unsigned>(o.done) << 30) | (pk & 0x80000000u);  // bit 31: canonical (county, column) pair
  return o;
}

// =========================================================================================
// step, variant THREAD: one thread per env.  The straightforward kernel: used for small
// batches (launch-latency bound anyway) and as the in-tree cross-check of the COOP kernel.
// =========================================================================================
template <int ADT>
__global__ void __launch_bounds__(128) step_thread_kernel(const StepArgs p) {
  const int stride = gridDim.x * blockDim.x;
  pdl_launch_dependents();
  pdl_wait();
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < p.n; e += stride) {
    const uint2 sv = ld_stream(p.st_static + e);
    uint2 mv = ld_stream(p.st_mut + e);
    const int a = load_action<ADT>(p.action, e);
    const Transition tr = advance(sv, mv, a);
    st_stream(p.st_mut + e, mv);

    const double* __restrict__ xr = p.tb.x + static_cast<size_t>(tr.row) * kXStride;
    const float* __restrict__ wv = p.tb.w + static_cast<size_t>(sv.y & kRecMask) * kWStride;
    float* ob = p.obs + static_cast<size_t>(e) * kObsDim;
    // bias first, then the 24 table features, then the three agent terms
    double zb = static_cast<double>(__ldg(wv + kWOther + 6)), ze = static_cast<double>(__ldg(wv + kWOther + 7));
    double heat_qi = 0.0;
#pragma unroll
    for (int u = 0; u < 24; ++u) {
      const double x = __ldg(xr + u);
      if (u == 0) heat_qi = x;
      zb = fma(x, static_cast<double>(__ldg(wv + 2 * u)), zb);
      ze = fma(x, static_cast<double>(__ldg(wv + 2 * u + 1)), ze);
      if (!tr.done) ob[slot_of_feature(u)] = static_cast<float>(x);  // obs of day t (Q2)
    }
    const double other[3] = {static_cast<double>(tr.lag), static_cast<double>(tr.streak),
                             static_cast<double>(tr.rem)};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      zb = fma(other[k], static_cast<double>(__ldg(wv + kWOther + 2 * k)), zb);
      ze = fma(other[k], static_cast<double>(__ldg(wv + kWOther + 2 * k + 1)), ze);
    }
    // reward = -1000/152 * sigmoid(z_b) * (1 - sigmoid(z_e) * [heat_qi > 0.5] * action)  (env.py:211-221)
    p.reward[e] = step_reward(zb, ze, tr.act && heat_qi > 0.5);
    p.terminated[e] = static_cast<uint8_t>(tr.done);
    if (p.remaining) p.remaining[e] = tr.rem;
    if (p.at_budget) p.at_budget[e] = static_cast<uint8_t>(tr.at_budget);
    if (!tr.done) {  // on and after the last day the previous observation stands (Q2)
      ob[23] = static_cast<float>(tr.lag);
      ob[24] = static_cast<float>(tr.streak);
      ob[25] = static_cast<float>(tr.rem);
      ob[27] = sig_to_obs(static_cast<float>(__ldg(xr + kXSig)));
      ob[28] = static_cast<float>(tr.cnt14);
    }
  }
}

// =========================================================================================
// Lane roles shared by the COOP step kernel and the reset kernel.  Lane q (0..7) of an
// env's 8-lane group loads record positions {2q, 2q+1} (load A) and {16+2q, 17+2q} (load
// B) of the day-row and the matching (b,e) coefficient pairs, and multiplies them.
// Positions 24..31 carry no table feature:
//   24  significance code x (0, 0)            -> contributes 0
//   25  heat_qi gate: 0.0 or -1e300 x (0, 1)  -> z_e stays, or becomes -1e300 and
//       sigmoid(z_e) is exactly 0, which is `effectiveness * (heat_qi > 0.5)` of
//       env.py:218 (quirk Q4) without a single instruction
//   26..31  zeros on the x side; the w side of 28..31 holds the non-table coefficients,
//       which lanes 6,7 hand to the env's owner thread through shared memory.
// For the observation a lane converts and stages the table slots it holds:
//   A.x -> 2q   A.y -> 2q+1   B.x -> 16+2q (q<4) | 27 significance (q=4)   B.y -> 17+2q (q<3) | 26 (q=3)
// =========================================================================================
struct LaneSlots {
  int bx, by;  // observation slot staged from B.x / B.y, -1 = none
};
__device__ __forceinline__ LaneSlots lane_slots(int q) {
  LaneSlots s;
  s.bx = q < 4 ? 16 + 2 * q : (q == 4 ? 27 : -1);
  s.by = q < 3 ? 17 + 2 * q : (q == 3 ? 26 : -1);
  return s;
}

#ifndef W2A_COOP_WARPS
#define W2A_COOP_WARPS 8  // warps per CTA
#endif
#ifndef W2A_COOP_MINBLOCKS
#define W2A_COOP_MINBLOCKS 2  // CTAs per SM the register allocation is sized for (B200 sweep: profiles/r01)
#endif
#ifndef W2A_COOP_AHEAD
#define W2A_COOP_AHEAD 3  // rounds of table loads in flight ahead of the one being consumed
#endif
#ifndef W2A_COOP_PRE
#define W2A_COOP_PRE 2  // rounds of the NEXT group whose loads are issued before this group's phase C
#endif
constexpr int kCoopWarps = W2A_COOP_WARPS;
constexpr int kAhead = W2A_COOP_AHEAD;
constexpr int kPre = W2A_COOP_PRE < W2A_COOP_AHEAD ? W2A_COOP_PRE : W2A_COOP_AHEAD;

// per-warp shared memory of the COOP kernel
struct __align__(128) WarpScratch {
  float obs[32 * kObsDim];  // the warp's 32 observation rows, same layout as in global memory (3712 B)
  float4 other[8][8];       // [round][2*group + (lane-6)]: raw non-table coefficients for the env's owner
  double2 part[8][32];      // [round][8*group + (lane%8 ^ swizzle)]: (pb, pe) partial dot products
};
constexpr int kCoopSmemBytes = kCoopWarps * static_cast<int>(sizeof(WarpScratch));
constexpr unsigned kObsBlockBytes = 32 * kObsDim * sizeof(float);

struct RoundLoads {
  double2 xa, xb;
  float4 wa, wb;
};

// L2 eviction policies: the tables are re-read every step and must survive the ~130 MB of
// per-step streaming traffic that flows through the 126 MB L2 (without the hints ncu shows the
// whole 19 MB coefficient table being re-fetched from HBM every launch).
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float4 ldg_keep(const float4* p, unsigned long long pol) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ double2 ldg_keep(const double2* p, unsigned long long pol) {
  double2 v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  return v;
}

// per-env state with an L2 evict_last policy: read and rewritten every step, so for env sets whose 16 B/env fit
// comfortably in L2 it never has to travel to HBM and back between two steps
__device__ __forceinline__ uint2 ld_keep_u2(const uint2* p, unsigned long long pol) {
  uint2 v;
  asm volatile("ld.global.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol) : "memory");
  return v;
}
__device__ __forceinline__ void st_keep_u2(uint2* p, uint2 v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.v2.u32 [%0], {%1,%2}, %3;" ::"l"(p), "r"(v.x), "r"(v.y), "l"(pol) : "memory");
}

__device__ __forceinline__ RoundLoads issue_round(const double2* x_lane, const float4* w_lane, unsigned row,
                                                  unsigned cv, unsigned long long keep) {
  RoundLoads r;
  const double2* xr = x_lane + static_cast<size_t>(row) * (kXStride / 2);
  const float4* wr = w_lane + static_cast<size_t>(cv) * (kWStride / 4);
  r.xa = __ldg(xr);  // the day's rows are hot within a step and dead afterwards: normal policy
  r.xb = __ldg(xr + 8);
  r.wa = ldg_keep(wr, keep);
  r.wb = ldg_keep(wr + 8, keep);
  return r;
}

// bulk (TMA) copy of a staged, 16-byte aligned shared-memory block to global memory
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, unsigned bytes, unsigned long long pol) {
  const unsigned saddr = static_cast<unsigned>(__cvta_generic_to_shared(ssrc));
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst), "r"(saddr),
               "r"(bytes), "l"(pol)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_all() {  // the writes themselves are done, not only the reads
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// =========================================================================================
// step, variant COOP: a warp owns 32 consecutive envs.
//
//   phase A  one thread per env: coalesced state/action traffic and the integer transition.
//   phase B  eight rounds, four envs per round, 8 lanes per env: four 128-bit loads per
//            lane fetch the env's day-row and coefficient record (the next round's loads
//            are issued before this round's math); the lane's float64 partial dot products
//            and the float32 copies of its observation slots go to shared memory.  The
//            round has no branch, no select and no integer work: the heat_qi gate is a
//            table column (see above) and the agent slots are the owner's business.
//   phase C  one thread per env: it sums the 8 partials of its env (XOR-swizzled, conflict
//            free), adds the three agent terms and the bias, evaluates the two float64
//            sigmoids, writes reward and flags, and patches its four agent slots into the
//            staged observation row; the warp's 32 rows (3712 contiguous bytes) then leave
//            through one bulk (TMA) store.
// =========================================================================================
// AHEAD = rounds of table loads in flight ahead of the one being consumed: kAhead for big batches (two CTAs
// per SM hide the rest), 7 = the whole group at once for small, launch-latency-bound batches.
template <int ADT, int AHEAD = kAhead>
__global__ void __launch_bounds__(32 * kCoopWarps, AHEAD > 4 ? 1 : W2A_COOP_MINBLOCKS) step_coop_kernel(const StepArgs p) {
  constexpr int PRE = kPre < AHEAD ? kPre : AHEAD;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  WarpScratch& ws = reinterpret_cast<WarpScratch*>(smem_raw)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int q = lane & 7;
  const int g = lane >> 3;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const int n_groups = (p.n + 31) >> 5;
  const double2* const x_lane = reinterpret_cast<const double2*>(p.tb.x) + q;
  const float4* const w_lane = reinterpret_cast<const float4*>(p.tb.w) + q;
  const LaneSlots ls = lane_slots(q);
  // Round m serves envs {2m, 2m+1, 2m+16, 2m+17} of the warp, group g taking env
  // row_of(m, g) = 2m + (g & 1) + 16 (g >> 1).  Two reasons for this composition:
  //   * neighbouring envs (2m, 2m+1) that share a day-row or a coefficient record in a
  //     locality-sorted env set fall into the same load instruction and share its lines;
  //   * with 29-float rows the four staged rows start at banks b, b+29, b+16, b+13 (mod 32):
  //     the 8 even-or-odd banks a group writes per store never collide with another group's.
  const int row_g = (g & 1) + 16 * (g >> 1);  // env of round 0 for this group; round m adds 2m
  float* const st_a = ws.obs + row_g * kObsDim + 2 * q;
  float* const st_bx = ws.obs + row_g * kObsDim + (ls.bx >= 0 ? ls.bx : 0);
  float* const st_by = ws.obs + row_g * kObsDim + (ls.by >= 0 ? ls.by : 0);
  // partials: slot [m][8g + (q ^ s)], s = 2*(m & 3) + (g & 1), so that the 8 owner threads of a
  // quarter-warp (four rounds x two groups) read 8 different bank quads
  double2* part_w[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) part_w[r] = &ws.part[0][8 * g + (q ^ (2 * r + (g & 1)))];  // round m: part_w[m & 3] + 32 m
  float4* const oth_w = &ws.other[0][2 * g + (q & 1)];              // lanes 6,7: + 8 m
  float* const my_row = ws.obs + lane * kObsDim;
  // the round and group that served env `lane`
  const int my_it = (lane & 15) >> 1, my_g = (lane & 1) + 2 * (lane >> 4);
  const int my_swz = 2 * (my_it & 3) + (my_g & 1);
  const bool obs_aligned = (reinterpret_cast<uintptr_t>(p.obs) & 15) == 0;
  bool bulk_pending = false;
  const unsigned long long keep = l2_policy_evict_last();        // table lines
  const unsigned long long stream_out = l2_policy_evict_first();  // observation rows

  // Software pipeline over the warp's groups of 32 envs: while group G is in phase C (fold,
  // sigmoids, stores), the state transition of group G+1 is already done and its first
  // AHEAD rounds of table loads are in flight; the per-env inputs of group G+2 are being
  // fetched.  `cur_*` belong to G, `nxt_*` to G+1, `in_*` are raw inputs one step further.
  int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  RoundLoads ring[AHEAD + 1];
  uint2 in_sv = make_uint2(0u, 0u), in_mv = make_uint2(0u, 0u);
  int in_a = 0;
  auto fetch_inputs = [&](int group) {
    if (group < n_groups) {
      const int ec = min((group << 5) + lane, p.n - 1);
      in_sv = ld_stream(p.st_static + ec);
      in_mv = ld_stream(p.st_mut + ec);
      in_a = load_action<ADT>(p.action, ec);
    }
  };
  Transition cur_tr = {};
  unsigned cur_cv = 0u;
  // phase A of `group` from the fetched inputs + its first AHEAD rounds of loads
  auto open_group = [&](int group) {
    if (group < n_groups) {
      const uint2 sv = in_sv;
      uint2 mv = in_mv;
      const int a = in_a;
      cur_tr = advance(sv, mv, a);
      const int e = (group << 5) + lane;
      if (e < p.n) st_stream(p.st_mut + e, mv);
      cur_cv = sv.y & kRecMask;
#pragma unroll
      for (int d = 0; d < PRE; ++d) {
        const unsigned rowd = __shfl_sync(kFull, cur_tr.row, 2 * d + row_g);
        const unsigned cvd = __shfl_sync(kFull, cur_cv, 2 * d + row_g);
        ring[d] = issue_round(x_lane, w_lane, rowd, cvd, keep);
      }
    }
  };
  pdl_launch_dependents();
  pdl_wait();  // the previous launch on the stream wrote the state (and the actions) this one reads
  fetch_inputs(wg);
  open_group(wg);
  fetch_inputs(wg + warps_total);

  for (; wg < n_groups; wg += warps_total) {
    const int base = wg << 5;
    const int e = base + lane;
    const bool valid = e < p.n;
    const Transition tr = cur_tr;
    const unsigned cv = cur_cv;
    const bool write_obs = valid && !tr.done;  // on and after the last day the previous obs stands (Q2)
    const unsigned wmask = __ballot_sync(kFull, write_obs);
    // the previous group's bulk store must have finished reading the staging buffer
    if (bulk_pending) {
      if (lane == 0) bulk_store_wait_read();
      __syncwarp();
      bulk_pending = false;
    }

    // ---- phase B ---------------------------------------------------------------------------
    // AHEAD rounds of loads are in flight while a round is being consumed; the first PRE of
    // them were issued one group earlier (open_group), the others follow now
#pragma unroll
    for (int d = PRE; d < AHEAD; ++d) {
      const unsigned rowd = __shfl_sync(kFull, tr.row, 2 * d + row_g);
      const unsigned cvd = __shfl_sync(kFull, cv, 2 * d + row_g);
      ring[d] = issue_round(x_lane, w_lane, rowd, cvd, keep);
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      if (it + AHEAD < 8) {
        const unsigned rowd = __shfl_sync(kFull, tr.row, 2 * (it + AHEAD) + row_g);
        const unsigned cvd = __shfl_sync(kFull, cv, 2 * (it + AHEAD) + row_g);
        ring[(it + AHEAD) % (AHEAD + 1)] = issue_round(x_lane, w_lane, rowd, cvd, keep);
      }
      const RoundLoads& cur = ring[it % (AHEAD + 1)];
      const double pb = fma(cur.xa.x, static_cast<double>(cur.wa.x), cur.xa.y * static_cast<double>(cur.wa.z)) +
                        fma(cur.xb.x, static_cast<double>(cur.wb.x), cur.xb.y * static_cast<double>(cur.wb.z));
      const double pe = fma(cur.xa.x, static_cast<double>(cur.wa.y), cur.xa.y * static_cast<double>(cur.wa.w)) +
                        fma(cur.xb.x, static_cast<double>(cur.wb.y), cur.xb.y * static_cast<double>(cur.wb.w));
      part_w[it & 3][it * 32] = make_double2(pb, pe);
      if (q >= 6) oth_w[it * 8] = cur.wb;
      // observation of day t (Q2): float32 of the table slots this lane holds
      float* const sa = st_a + it * (2 * kObsDim);
      sa[0] = static_cast<float>(cur.xa.x);
      sa[1] = static_cast<float>(cur.xa.y);
      if (ls.bx >= 0) st_bx[it * (2 * kObsDim)] = static_cast<float>(cur.xb.x);
      if (ls.by >= 0) st_by[it * (2 * kObsDim)] = static_cast<float>(cur.xb.y);
    }
    __syncwarp();

    // ---- open the next group: its loads fly while this one finishes --------------------------------
    open_group(wg + warps_total);
    fetch_inputs(wg + 2 * warps_total);

    // ---- phase C ---------------------------------------------------------------------------
    // env (lane) was served in round my_it by group my_g: partials in part[my_it][8*my_g + (j ^ my_swz)],
    // its non-table coefficients in other[my_it][2*my_g], other[my_it][2*my_g + 1]
    {
      const float4 o0 = ws.other[my_it][2 * my_g];      // (b,e) of alert_lag1, alert_streak
      const float4 o1 = ws.other[my_it][2 * my_g + 1];  // (b,e) of remaining_budget, bias
      double zb = static_cast<double>(o1.z), ze = static_cast<double>(o1.w);
      const double f_lag = static_cast<double>(tr.lag), f_streak = static_cast<double>(tr.streak),
                   f_rem = static_cast<double>(tr.rem);
      zb = fma(f_lag, static_cast<double>(o0.x), zb);
      ze = fma(f_lag, static_cast<double>(o0.y), ze);
      zb = fma(f_streak, static_cast<double>(o0.z), zb);
      ze = fma(f_streak, static_cast<double>(o0.w), ze);
      zb = fma(f_rem, static_cast<double>(o1.x), zb);
      ze = fma(f_rem, static_cast<double>(o1.y), ze);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const double2 v = ws.part[my_it][8 * my_g + (j ^ my_swz)];
        zb += v.x;
        ze += v.y;
      }
      // the four agent slots and the significance NaN of this env's staged observation row
      my_row[23] = static_cast<float>(tr.lag);
      my_row[24] = static_cast<float>(tr.streak);
      my_row[25] = static_cast<float>(tr.rem);
      my_row[27] = sig_to_obs(my_row[27]);
      my_row[28] = static_cast<float>(tr.cnt14);
      if (valid) {
        __stcs(p.reward + e, step_reward(zb, ze, tr.act));
        p.terminated[e] = static_cast<uint8_t>(tr.done);
        if (p.remaining) __stcs(p.remaining + e, tr.rem);
        if (p.at_budget) p.at_budget[e] = static_cast<uint8_t>(tr.at_budget);
      }
    }
    fence_proxy_async_smem();  // staged observation rows -> visible to the bulk-copy engine
    __syncwarp();
    // observation rows: 32 x 116 B contiguous in global memory (base is a multiple of 32)
    float* const ob = p.obs + static_cast<size_t>(base) * kObsDim;
    if (wmask == kFull && obs_aligned) {
      if (lane == 0) bulk_store(ob, ws.obs, kObsBlockBytes, stream_out);
      bulk_pending = true;
    } else if (wmask != 0u) {  // ragged tail or some envs already on their last day
      for (int i = lane; i < 32 * kObsDim; i += 32) {
        const int env = (i * 565) >> 14;  // i / 29 for i < 928
        if ((wmask >> env) & 1u) st_stream(ob + i, ws.obs[i]);
      }
    }
    __syncwarp();
  }
  if (bulk_pending && lane == 0) bulk_store_wait_all();
}

// =========================================================================================
// step, variant TABLE: one 32-byte sector + one 128-byte observation row per env.
//
// The table part of both logits depends on (county, year, day, posterior sample) only -- never on
// the agent -- so it is evaluated ONCE for the whole feature store when the tables are installed
// (w2a_precompute_logits: the reference's per-key loop env.py:206-217 hoisted out of `step`):
//   logit_all_kernel   (z_b, z_e) of every (day, column, sample, year), gate and bias included -- a float64
//                      GEMM per (county, day) on the FP64 tensor path (DMMA) -- written as z32_all (one
//                      sector per entry, with the agent coefficients) and zt_all (float64, episode-major)
//   obs_rows_kernel    x32[row][0..28] = the float32 observation row of every day-row
//   step_table_kernel  per env: transition, z = its entry of ITS OWN day + agent terms, sigmoids,
//                      reward (one gathered sector); per warp: the 32 observation rows are copies of
//                      x32 rows (one gathered line per env), patched with the agent slots, and
//                      leave through one bulk (TMA) store.
// Every env looks its own day up, so envs need not march in lock step.  An env whose coefficient
// column is not its own county (augmentation, quirk Q6: no `canonical` bit) computes its logits
// directly from x / w (the THREAD kernel's arithmetic): results never depend on the variant.
// =========================================================================================

#ifndef W2A_TABLE_TASK_WARPS
#define W2A_TABLE_TASK_WARPS 4
#endif
constexpr int kTableTaskWarps = W2A_TABLE_TASK_WARPS;  // warps (= tasks) per CTA of logit_all_kernel
#ifndef W2A_TABLE_BUILD_MINBLOCKS
#define W2A_TABLE_BUILD_MINBLOCKS 4  // x 4 warps; a one-off kernel, sized so that nothing spills
#endif

// D(8x8) += A(8x4) * B(4x8) in float64 on the tensor cores (DMMA).  Fragments: A[lane/4][lane%4],
// B[lane%4][lane/4], C/D[lane/4][2*(lane%4) + {0,1}].
__device__ __forceinline__ void dmma_884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// The logit table is a small float64 GEMM per (county, day) -- Z[(p, head), y] = W[(p, head), 0..23] .
// X[y, 0..23]^T, plus the bias and (for z_e) the gate column -- so it runs on the FP64 tensor path:
// one warp per task = (day, coefficient column, block of 8 posterior samples): 6 k-steps x 2 heads
// per tile of 8 years of m8n8k4.  Everything that depends on the task index alone
// (coefficients, line heads) is requested before the first wait, the day-rows of the first two year
// tiles right after.
__global__ void __launch_bounds__(32 * kTableTaskWarps, W2A_TABLE_BUILD_MINBLOCKS)
    logit_all_kernel(const Tables tb, double2* zt_all, float4* z32_all, int t0) {
  const int lane = threadIdx.x & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const int mtiles = (tb.P + 7) >> 3;
  // blockIdx.x = day: the days of one task are scheduled back to back, so their 16-byte zt_all entries (8 days per
  // line) meet in L2 before the line is written back
  const int task = blockIdx.y * kTableTaskWarps + (threadIdx.x >> 5);
  const int t = t0 + static_cast<int>(blockIdx.x);
  if (task >= tb.Sw * mtiles || t >= tb.T) return;
  const int col = task / mtiles, mt = task - col * mtiles;
  const int s = __ldg(tb.col_to_loc + col);
  const int p = mt * 8 + gid;  // this lane's row of A and C
  const bool p_ok = p < tb.P;
  const float2* wrec = reinterpret_cast<const float2*>(tb.w + static_cast<size_t>(col * tb.P + min(p, tb.P - 1)) * kWStride);
  float2 wv[6];  // A fragments of the two heads: coefficient pair of record position 4*ks + tig
#pragma unroll
  for (int ks = 0; ks < 6; ++ks) wv[ks] = __ldg(wrec + 4 * ks + tig);
  const float2 bias = __ldg(wrec + 31);
  const float4 c_ls = __ldg(reinterpret_cast<const float4*>(wrec) + 14);     // (b, e) of alert_lag1, alert_streak
  const float2 c_rem = __ldg(wrec + 30);                                     // (b, e) of remaining_budget
  if (s < 0) return;  // a column without a county in the store: its lines stay zero, no env is canonical on it
  float4* rec32 = z32_all + ((static_cast<size_t>(t) * tb.Sw * tb.P + static_cast<size_t>(col * tb.P + p)) * tb.Y) * 2;
  double2* rect = zt_all + (static_cast<size_t>(col * tb.P + p) * tb.Y) * tb.T + t;  // + year * T
  const double* xday = tb.x + (static_cast<size_t>(s) * tb.Y * tb.T + t) * kXStride;
  const size_t ystride = static_cast<size_t>(tb.T) * kXStride;
  // year tiles of 8: B fragments x[y0 + gid][4*ks + tig] and, for this lane's two columns of C,
  // the gate x[y][25] (coefficients (0, 1): added to z_e as it is)
  auto load_b = [&](double (&xb)[6], double (&gate)[2], int y0) {
    const int yb = min(y0 + gid, tb.Y - 1);  // columns past Y are computed on a copy of the last year, never stored
#pragma unroll
    for (int ks = 0; ks < 6; ++ks) xb[ks] = __ldg(xday + yb * ystride + 4 * ks + tig);
#pragma unroll
    for (int i = 0; i < 2; ++i) gate[i] = __ldg(xday + min(y0 + 2 * tig + i, tb.Y - 1) * ystride + 25);
  };
  double xb0[6], xb1[6], g0[2], g1[2];
  load_b(xb0, g0, 0);
  load_b(xb1, g1, 8);
  double ab[6], ae[6];
#pragma unroll
  for (int ks = 0; ks < 6; ++ks) {
    ab[ks] = wv[ks].x;
    ae[ks] = wv[ks].y;
  }
  auto tile = [&](const double (&xb)[6], const double (&gate)[2], int y0) {
    double cb[2] = {bias.x, bias.x}, ce[2] = {bias.y + gate[0], bias.y + gate[1]};
#pragma unroll
    for (int ks = 0; ks < 6; ++ks) {
      dmma_884(cb, ab[ks], xb[ks]);
      dmma_884(ce, ae[ks], xb[ks]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int y = y0 + 2 * tig + i;  // this lane's columns of C
      if (p_ok && y < tb.Y) {
        rect[static_cast<size_t>(y) * tb.T] = make_double2(cb[i], ce[i]);
        rec32[2 * y] = c_ls;
        rec32[2 * y + 1] = make_float4(c_rem.x, c_rem.y, static_cast<float>(cb[i]), static_cast<float>(ce[i]));
      }
    }
  };
  tile(xb0, g0, 0);
  if (tb.Y > 8) tile(xb1, g1, 8);
  for (int y0 = 16; y0 < tb.Y; y0 += 8) {
    load_b(xb0, g0, y0);
    tile(xb0, g0, y0);
  }
}

// float32 observation rows (slot order) of every day-row: one warp per row, lane = slot.  Slots 23..25
// and 28 (agent slots, patched by the step kernel) and the pad lanes 29..31 are zero.
__global__ void __launch_bounds__(256) obs_rows_kernel(const double* __restrict__ x, float* __restrict__ x32, unsigned rows) {
  const int lane = threadIdx.x & 31;
  const unsigned warps_total = (gridDim.x * blockDim.x) >> 5;
  const int pos = lane < 23 ? lane : lane == 26 ? 23 : lane == 27 ? kXSig : -1;
  for (unsigned r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps_total) {
    float f = pos >= 0 ? static_cast<float>(__ldg(x + static_cast<size_t>(r) * kXStride + pos)) : 0.f;
    if (lane == 27) f = sig_to_obs(f);
    x32[static_cast<size_t>(r) * 32 + lane] = f;
  }
}

#ifndef W2A_Z_LOAD
#define W2A_Z_LOAD 1  // how the env's 32-byte sector is fetched: 1 one 256-bit load per env (LDG.E.256), 0 two 128-bit loads
#endif
struct __align__(128) TableScratchWarp {
  float obs[32 * kObsDim];
};
#ifndef W2A_TABLE_WARPS
#define W2A_TABLE_WARPS 4  // warps per CTA of step_table_kernel (B200 sweep: 2..16 within 2 %)
#endif
constexpr int kTableWarps = W2A_TABLE_WARPS;
constexpr int kTableSmemBytes = kTableWarps * static_cast<int>(sizeof(TableScratchWarp));

#ifndef W2A_TABLE_ROWS
#define W2A_TABLE_ROWS 2  // rounds of observation-row gathers in flight per warp (B200 sweep: 2 with 4 CTAs/SM >= 8 with 3)
#endif
#ifndef W2A_TABLE_MINBLOCKS
#define W2A_TABLE_MINBLOCKS 8  // x 4 warps = 32 warps/SM at 64 registers
#endif
#ifndef W2A_Z_POLICY
#define W2A_Z_POLICY 0  // L2 policy of the sector gather (two 128-bit loads only): 0 default, 1 evict_last
#endif
#ifndef W2A_STATE_KEEP
#define W2A_STATE_KEEP 1  // per-env state loads/stores carry an L2 evict_last policy when StepArgs::keep_state is set
#endif
#ifndef W2A_KNOCK
#define W2A_KNOCK 0  // timing studies only (results become wrong; profiles/r02/knockouts.txt): 1 no observation store,
#endif               // 2 sector gather from entry 0, 4 row gather from row 0, 8 no reward / flag outputs, 16 sectors of day 0
                     // whatever the day (the slice stays in L2), 64 / 128 the sector gather with 4 / 1 distinct lines per load

// ROWS = rounds of row gathers in flight per warp.  2 (W2A_TABLE_ROWS) where the step is bandwidth-bound: env slots with
// locality, or more than ~5M envs; 1 for up to ~5M envs without locality, where the cold sector gather sets the pace and a
// second row round in flight only gets in its way (B200, 2^20 interleaved envs: 33.3 vs 34.5 us; sorted: 29.7 vs 28.5;
// 2^23: 248 vs 238); 8 (every gather of the group issued before anything waits) for env sets of a few groups per SM,
// where a step is a chain of memory latencies.
// The one-round instance runs 7 CTAs per SM (72 registers) instead of 8: 33.0 vs 33.3 us at 2^20 envs, 17.7 vs 19.4 at 2^19,
// 6.4 vs 6.8 at 2^17 -- fewer warps asking at once, again; the two-round instances lose 3 % with 7.
// LATE = no row gather is issued before the group's logits and rewards are done (instead of the first ROWS rounds right
// behind the sector): for slots in table order whose state stays in L2 (sorted 2^20 envs: 27.2 vs 28.5 us; no difference
// without locality; slower when the state streams from HBM: 2^23 envs 249 vs 238).
// `post` hooks a consumer of the fresh observations into the kernel (w2a_policy.cuh: the MLP policy that picks the next
// action from the rows while they are still in shared memory); StepNoPost is the plain step.
struct StepNoPost {
  __device__ __forceinline__ void reward(int, float) const {}
  __device__ __forceinline__ void group(const float*, int, int) const {}
};
template <int ADT, int ROWS, bool LATE, typename Post>
__device__ __forceinline__ void step_table_body(const StepArgs& p, Post& post) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* const stage = reinterpret_cast<TableScratchWarp*>(smem_raw)[threadIdx.x >> 5].obs;
  const int lane = threadIdx.x & 31;
  const int q = lane & 7;
  const int g = lane >> 3;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const int n_groups = (p.n + 31) >> 5;
#if W2A_Z_POLICY == 1
  const unsigned long long keep = l2_policy_evict_last();
#endif
  const unsigned long long stream_out = l2_policy_evict_first();
  const bool obs_aligned = (reinterpret_cast<uintptr_t>(p.obs) & 15) == 0;
  const float4* const x32_lane = reinterpret_cast<const float4*>(p.tb.x32) + q;
  float* const my_row = stage + lane * kObsDim;
  bool bulk_pending = false;

  pdl_launch_dependents();
  int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wg >= n_groups) return;
  pdl_wait();  // the previous launch on the stream wrote the state (and the actions) this one reads
  // the next group's state and action are in flight while this group's gathers are waited for
  uint2 sv_n, mv_n;
  int a_n;
#if W2A_STATE_KEEP
  const unsigned long long keep_state = l2_policy_evict_last();
#endif
  auto fetch = [&](int group) {
    const int ec = min((group << 5) + lane, p.n - 1);
#if W2A_STATE_KEEP
    if (p.keep_state) {
      sv_n = ld_keep_u2(p.st_static + ec, keep_state);
      mv_n = ld_keep_u2(p.st_mut + ec, keep_state);
    } else
#endif
    {
      sv_n = ld_stream(p.st_static + ec);
      mv_n = ld_stream(p.st_mut + ec);
    }
    a_n = load_action<ADT>(p.action, ec);
  };
  fetch(wg);
  for (; wg < n_groups; wg += warps_total) {
    const int base = wg << 5;
    const int e = base + lane;
    const bool valid = e < p.n;
    const uint2 sv = sv_n;
    uint2 mv = mv_n;
    const int a = a_n;
    const unsigned t_before = mv.y & 0xffu;
    const bool fast = (mv.y >> 31) != 0;  // canonical: the env's column is its own county -> its logits are tabulated
    const Transition tr = advance(sv, mv, a);
    const unsigned cv = sv.y & kRecMask;
    const unsigned yr = sv.y >> 24;
    // ---- every gather of the group is issued before anything waits on one ---------------------------
    // (loads are unconditional -- an env off the fast path reads line 0 and ignores it -- so that no
    // select sits between a load and the next one)
    // the env's sector of z32_all: agent coefficients + float32 (z_b, z_e) of its own (day, column, sample, year)
    const unsigned entry = (fast && !(W2A_KNOCK & 2))
        ? (((W2A_KNOCK & 16) ? 0u : t_before) * p.tb.n_records + cv) * static_cast<unsigned>(p.tb.Y) + yr
        : 0u;
#if W2A_KNOCK & 64   // timing study: the 8 lanes of a group read the group leader's sector (4 distinct lines per load)
    const float4* zrec = p.tb.z32_all + static_cast<size_t>(__shfl_sync(kFull, entry, lane & 24)) * 2;
#elif W2A_KNOCK & 128  // timing study: the whole warp reads lane 0's sector (1 line per load)
    const float4* zrec = p.tb.z32_all + static_cast<size_t>(__shfl_sync(kFull, entry, 0)) * 2;
#else
    const float4* zrec = p.tb.z32_all + static_cast<size_t>(entry) * 2;
#endif
#if W2A_Z_LOAD == 1  // one 256-bit load of the sector
    float4 o0, o1;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(o0.x), "=f"(o0.y), "=f"(o0.z), "=f"(o0.w), "=f"(o1.x), "=f"(o1.y), "=f"(o1.z), "=f"(o1.w)
                 : "l"(zrec));
#elif W2A_Z_POLICY == 1
    const float4 o0 = ldg_keep(zrec, keep), o1 = __ldg(zrec + 1);
#else
    const float4 o0 = __ldg(zrec), o1 = __ldg(zrec + 1);
#endif
    float4 rows[ROWS];  // observation rows: 8 lanes copy one float32 row, four neighbouring envs per round
    auto load_rows = [&](int first) {
#pragma unroll
      for (int it = 0; it < ROWS; ++it) {
        const int j = 4 * (first + it) + g;
        const unsigned rowj = (W2A_KNOCK & 4) ? 0u : __shfl_sync(kFull, tr.row, j);
        rows[it] = __ldg(x32_lane + static_cast<size_t>(rowj) * 8);
      }
    };
    auto stage_rows = [&](int first) {
#pragma unroll
      for (int it = 0; it < ROWS; ++it) {
        const int j = 4 * (first + it) + g;
        float* st = stage + j * kObsDim + 4 * q;
        st[0] = rows[it].x;
        if (q < 7) {
          st[1] = rows[it].y;
          st[2] = rows[it].z;
          st[3] = rows[it].w;
        }
      }
    };
    if (!LATE) load_rows(0);
#if W2A_STATE_KEEP
    if (valid && p.keep_state) st_keep_u2(p.st_mut + e, mv, keep_state);
    else
#endif
    if (valid) st_stream(p.st_mut + e, mv);
    if (wg + warps_total < n_groups) fetch(wg + warps_total);
    const bool write_obs = valid && !tr.done;
    const unsigned wmask = __ballot_sync(kFull, write_obs);
    if (bulk_pending) {  // the previous group's bulk store must have finished reading the staging buffer
      if (lane == 0) bulk_store_wait_read();
      __syncwarp();
      bulk_pending = false;
    }
    // ---- logits -------------------------------------------------------------------------------
    double zb = static_cast<double>(o1.z), ze = static_cast<double>(o1.w);
    float lag_b = o0.x, lag_e = o0.y, streak_b = o0.z, streak_e = o0.w, rem_b = o1.x, rem_e = o1.y;
    if (!fast) {  // non-canonical env (similar-county column, Q6): straight from the tables
      const double* xr = p.tb.x + static_cast<size_t>(tr.row) * kXStride;
      const float* wv = p.tb.w + static_cast<size_t>(cv) * kWStride;
      zb = __ldg(xr + 25) * static_cast<double>(__ldg(wv + 50));
      ze = __ldg(xr + 25) * static_cast<double>(__ldg(wv + 51));
      for (int u = 0; u < 24; ++u) {
        const double x = __ldg(xr + u);
        zb = fma(x, static_cast<double>(__ldg(wv + 2 * u)), zb);
        ze = fma(x, static_cast<double>(__ldg(wv + 2 * u + 1)), ze);
      }
      zb += static_cast<double>(__ldg(wv + kWOther + 6));
      ze += static_cast<double>(__ldg(wv + kWOther + 7));
      lag_b = __ldg(wv + kWOther); lag_e = __ldg(wv + kWOther + 1);
      streak_b = __ldg(wv + kWOther + 2); streak_e = __ldg(wv + kWOther + 3);
      rem_b = __ldg(wv + kWOther + 4); rem_e = __ldg(wv + kWOther + 5);
    }
    {
      const double f_lag = static_cast<double>(tr.lag), f_streak = static_cast<double>(tr.streak),
                   f_rem = static_cast<double>(tr.rem);
      double ab = f_lag * static_cast<double>(lag_b), ae = f_lag * static_cast<double>(lag_e);
      ab = fma(f_streak, static_cast<double>(streak_b), ab);
      ae = fma(f_streak, static_cast<double>(streak_e), ae);
      ab = fma(f_rem, static_cast<double>(rem_b), ab);
      ae = fma(f_rem, static_cast<double>(rem_e), ae);
      // The tabulated z is a float32: exact to 1e-6 below |z| = 32.  A bigger |z| only matters if the agent terms
      // pull the logit back into the range where the sigmoid is not saturated (|z + a| < 30; beyond it the reward
      // changes by less than 1e-13 relative or is below 1e-12 absolute): then -- and for magnitudes where float32
      // cannot even tell -- take the float64 pair of zt_all.  (+-inf is the closed heat_qi gate: exact as it is.)
      const float mb = fabsf(o1.z), me = fabsf(o1.w);
      const bool need_b = mb >= 32.f && mb <= 3.0e38f && (fabs(zb + ab) < 30.0 || mb > 1.0e6f);
      const bool need_e = me >= 32.f && me <= 3.0e38f && (fabs(ze + ae) < 30.0 || me > 1.0e6f);
      if (fast && (need_b || need_e)) {
        const double2 z = __ldg(p.tb.zt_all + (static_cast<size_t>(cv) * p.tb.Y + yr) * p.tb.T + t_before);
        zb = z.x;
        ze = z.y;
      }
      zb += ab;
      ze += ae;
    }
    if (valid && !(W2A_KNOCK & 8)) {
      const float r = step_reward(zb, ze, tr.act);
      __stcs(p.reward + e, r);
      post.reward(e, r);
      p.terminated[e] = static_cast<uint8_t>(tr.done);
      if (p.remaining) __stcs(p.remaining + e, tr.rem);
      if (p.at_budget) p.at_budget[e] = static_cast<uint8_t>(tr.at_budget);
    } else if (W2A_KNOCK & 8) {
      if (step_reward(zb, ze, tr.act) == 123.f) p.reward[e] = 0.f;  // keeps the arithmetic alive
    }
#pragma unroll
    for (int first = 0; first < 8; first += ROWS) {
      if (first > 0 || LATE) load_rows(first);
      stage_rows(first);
    }
    __syncwarp();
    my_row[23] = static_cast<float>(tr.lag);
    my_row[24] = static_cast<float>(tr.streak);
    my_row[25] = static_cast<float>(tr.rem);
    my_row[28] = static_cast<float>(tr.cnt14);
    fence_proxy_async_smem();
    __syncwarp();
    float* const ob = p.obs + static_cast<size_t>(base) * kObsDim;
    if (W2A_KNOCK & 1) {
      if (stage[lane] == 123.f) ob[lane] = 0.f;
    } else if (wmask == kFull && obs_aligned) {
      if (lane == 0) bulk_store(ob, stage, kObsBlockBytes, stream_out);
      bulk_pending = true;
    } else if (wmask != 0u) {
      for (int i = lane; i < 32 * kObsDim; i += 32) {
        const int env = (i * 565) >> 14;  // i / 29 for i < 928
        if ((wmask >> env) & 1u) st_stream(ob + i, stage[i]);
      }
    }
    post.group(stage, base, min(32, p.n - base));  // the group's 32 rows as the observation tensor holds them
    __syncwarp();
  }
  if (bulk_pending && lane == 0) bulk_store_wait_all();
}

template <int ADT, int ROWS = W2A_TABLE_ROWS, bool LATE = false>
__global__ void __launch_bounds__(32 * kTableWarps, ROWS == 1 ? W2A_TABLE_MINBLOCKS - 1 : ROWS <= 4 ? W2A_TABLE_MINBLOCKS : 4)
    step_table_kernel(const StepArgs p) {
  StepNoPost post;
  step_table_body<ADT, ROWS, LATE>(p, post);
}

// =========================================================================================
// reset: host index streams -> state + first observation (env.py:155-184).
// Same warp shape and lane roles as the COOP step; observations go straight to global
// memory (a reset happens once per 153 steps).
// =========================================================================================
__global__ void __launch_bounds__(256) reset_kernel(const ResetArgs p) {
  const int lane = threadIdx.x & 31;
  const int q = lane & 7;
  const LaneSlots ls = lane_slots(q);
  const double2* const x_lane = reinterpret_cast<const double2*>(p.tb.x) + q;
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
    // canonical pair: the coefficient column is the episode's own county (no similar-county
    // substitution, Q6), so the per-step logit table applies to this env
    const int colv = __ldcs(p.coef_col + ec);
    const bool canon = p.tb.col_to_loc != nullptr && __ldg(p.tb.col_to_loc + colv) == __ldcs(p.loc + ec);
    if (take) {
      st_stream(p.st_static + e, make_uint2(row0 | (nd << 24), cv | (static_cast<unsigned>(__ldcs(p.year + ec)) << 24)));
      st_stream(p.st_mut + e, make_uint2(static_cast<unsigned>(budget), canon ? 0x80000000u : 0u));  // t = streak = win = 0
      if (p.remaining) p.remaining[e] = budget;
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int j = 2 * it + ((lane >> 3) & 1) + 16 * (lane >> 4);  // env served by this lane's group in this round
      const unsigned rowj = __shfl_sync(kFull, row0, j);
      const int takej = __shfl_sync(kFull, static_cast<int>(take), j);
      if (takej) {
        const double2* xr = x_lane + static_cast<size_t>(rowj) * (kXStride / 2);
        const double2 xa = __ldg(xr), xb = __ldg(xr + 8);
        float vbx = static_cast<float>(xb.x);
        if (q == 4) vbx = sig_to_obs(vbx);
        float* ob = p.obs + static_cast<size_t>(base + j) * kObsDim;
        st_stream(ob + 2 * q, static_cast<float>(xa.x));
        st_stream(ob + 2 * q + 1, static_cast<float>(xa.y));
        if (ls.bx >= 0) st_stream(ob + ls.bx, vbx);
        if (ls.by >= 0) st_stream(ob + ls.by, static_cast<float>(xb.y));
      }
    }
    if (take) {  // agent slots of the first observation (env.py:190-193 at t = 0)
      float* ob = p.obs + static_cast<size_t>(e) * kObsDim;
      st_stream(ob + 23, 0.f);                          // alert_lag1
      st_stream(ob + 24, 0.f);                          // alert_streak
      st_stream(ob + 25, static_cast<float>(budget));   // remaining_budget = budget
      st_stream(ob + 28, 0.f);                          // alert_2wks
    }
  }
}

// The same reset once the float32 observation rows exist (w2a_precompute_logits): the first observation is the row of
// day 0 as it lies in x32 plus the remaining budget -- one 128-byte gather per env (8 lanes per row) instead of a
// 256-byte float64 record and 29 scalar stores, staged and written with one bulk store per full group like the TABLE step.
__global__ void __launch_bounds__(32 * kTableWarps, 8) reset_table_kernel(const ResetArgs p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* const stage = reinterpret_cast<TableScratchWarp*>(smem_raw)[threadIdx.x >> 5].obs;
  const int lane = threadIdx.x & 31;
  const int q = lane & 7;
  const int g = lane >> 3;
  const float4* const x32_lane = reinterpret_cast<const float4*>(p.tb.x32) + q;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const int n_groups = (p.n + 31) >> 5;
  const unsigned long long stream_out = l2_policy_evict_first();
  const bool obs_aligned = (reinterpret_cast<uintptr_t>(p.obs) & 15) == 0;
  bool bulk_pending = false;
  for (int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; wg < n_groups; wg += warps_total) {
    const int base = wg << 5;
    const int e = base + lane;
    const bool valid = e < p.n;
    const int ec = valid ? e : p.n - 1;
    const bool take = valid && (p.mask == nullptr || p.mask[ec] != 0);
    const int locv = __ldcs(p.loc + ec), yearv = __ldcs(p.year + ec), colv = __ldcs(p.coef_col + ec);
    const int sy = locv * p.tb.Y + yearv;
    const unsigned nd = __ldg(p.tb.n_days + sy);
    const unsigned row0 = static_cast<unsigned>(sy) * static_cast<unsigned>(p.tb.T);
    int budget = p.budget ? __ldcs(p.budget + ec) : INT32_MIN;
    if (budget == INT32_MIN) budget = __ldg(p.tb.budget0 + sy);  // env.py:167-170
    const unsigned cv = static_cast<unsigned>(colv) * static_cast<unsigned>(p.tb.P) + static_cast<unsigned>(__ldcs(p.coef_idx + ec));
    const bool canon = p.tb.col_to_loc != nullptr && __ldg(p.tb.col_to_loc + colv) == locv;
    float4 rows[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) rows[it] = __ldg(x32_lane + static_cast<size_t>(__shfl_sync(kFull, row0, 4 * it + g)) * 8);
    if (take) {
      st_stream(p.st_static + e, make_uint2(row0 | (nd << 24), cv | (static_cast<unsigned>(yearv) << 24)));
      st_stream(p.st_mut + e, make_uint2(static_cast<unsigned>(budget), canon ? 0x80000000u : 0u));  // t = streak = win = 0
      if (p.remaining) p.remaining[e] = budget;
    }
    const unsigned wmask = __ballot_sync(kFull, take);
    if (bulk_pending) {  // the previous group's bulk store must have finished reading the staging buffer
      if (lane == 0) bulk_store_wait_read();
      __syncwarp();
      bulk_pending = false;
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      float* st = stage + (4 * it + g) * kObsDim + 4 * q;
      st[0] = rows[it].x;
      if (q < 7) {
        st[1] = rows[it].y;
        st[2] = rows[it].z;
        st[3] = rows[it].w;
      }
    }
    __syncwarp();
    stage[lane * kObsDim + 25] = static_cast<float>(budget);  // remaining_budget = budget; lag, streak, alerts_2wks are 0 in x32
    fence_proxy_async_smem();
    __syncwarp();
    float* const ob = p.obs + static_cast<size_t>(base) * kObsDim;
    if (wmask == kFull && obs_aligned) {
      if (lane == 0) bulk_store(ob, stage, kObsBlockBytes, stream_out);
      bulk_pending = true;
    } else if (wmask != 0u) {  // masked reset or ragged tail: the other envs keep their observation
      for (int i = lane; i < 32 * kObsDim; i += 32) {
        const int env = (i * 565) >> 14;  // i / 29 for i < 928
        if ((wmask >> env) & 1u) st_stream(ob + i, stage[i]);
      }
    }
    __syncwarp();
  }
  if (bulk_pending && lane == 0) bulk_store_wait_all();
}

// =========================================================================================
// rollout: the whole `while not done: step(a)` loop (env.py:273-276) of every env in one
// launch.  Observations are not materialised; what comes out is the sum of rewards per env,
// the advanced state, and the sufficient statistics of the reference's logging callback
// (callbacks.py:63-74).
//
// 8 lanes per env, 4 envs per warp, and time is processed in chunks of 8 days:
//   * the env's coefficient record is loaded and converted to float64 ONCE (registers);
//   * per day, lane q loads record positions {2q,2q+1} and {16+2q,17+2q} of the day-row
//     (the env's days are consecutive 256-byte records) and accumulates its partial logits;
//   * a transposing shuffle butterfly hands day (chunk, q) to lane q: the 8 lanes of an env
//     evaluate 8 different days' sigmoids at once;
//   * the integer recurrence of the chunk (budget clamp, streak, 14-day window) depends on
//     actions only, never on rewards, so every lane replays it from the chunk's 8 action
//     bits and picks the values of its own day.
// Actions come from a [steps, n] byte bank or from a counter-based Bernoulli(p) hash.
// An env that is already done is left alone (unlike `step`, quirk Q11).
// =========================================================================================
// pt[row] = the logistic policy's logit over the 24 table slots of day-row `row` (the agent-independent part of
// `policy(obs)`): one thread per row, once per w2a_rollout_policy call (the policy changes from call to call)
__global__ void __launch_bounds__(256) policy_rows_kernel(const double* __restrict__ x, const double* __restrict__ policy,
                                                          double* __restrict__ pt, unsigned rows) {
  __shared__ double th[24];
  if (threadIdx.x < 24) th[threadIdx.x] = policy[slot_of_feature(threadIdx.x)];
  __syncthreads();
  for (unsigned r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
    const double2* xr = reinterpret_cast<const double2*>(x + static_cast<size_t>(r) * kXStride);
    double acc = 0.0;
#pragma unroll
    for (int u = 0; u < 12; ++u) {
      const double2 v = __ldg(xr + u);
      acc = fma(v.x, th[2 * u], acc);
      acc = fma(v.y, th[2 * u + 1], acc);
    }
    pt[r] = acc;
  }
}

struct RolloutArgs {
  Tables tb;
  const uint2* st_static;
  uint2* st_mut;
  const uint8_t* actions;  // [steps, n] or null
  unsigned long long seed;
  unsigned threshold;      // P(action = 1) * 2^32 when actions == null
  double* ep_return;       // [n], +=
  double* stats;           // [kNumStats], += (see StatIndex)
  int n, steps;
  // closed-loop logistic policy (rollout_kernel<true>): kPolicyLen float64 = one weight per observation
  // slot (info["feature_names"] order; slot 27, the significance label, is ignored) and the bias
  const double* policy;
  int stochastic;          // 0: action = logit > 0; 1: action = hash(seed, env, step) < sigmoid(logit) * 2^32
  int only;                // 0 every env, 1 canonical envs only (the table launch), 2 non-canonical envs only (its direct complement)
};
constexpr int kPolicyLen = kObsDim + 1;

// One butterfly level of the transposing reduction: lanes whose bit is set keep the upper
// half of the values and send the lower half, and vice versa; after levels 4, 2, 1 lane q of
// an 8-lane group holds the group's sum of value q.
template <int N>
__device__ __forceinline__ void fold_level(double (&v)[N], bool upper, int lane_xor) {
#pragma unroll
  for (int k = 0; k < N / 2; ++k) {
    const double keep = upper ? v[k + N / 2] : v[k];
    const double send = upper ? v[k] : v[k + N / 2];
    v[k] = keep + __shfl_xor_sync(kFull, send, lane_xor);
  }
}

// counter-based action stream: bit = hash(seed, env, step) < threshold   (splitmix64 finaliser)
__host__ __device__ __forceinline__ unsigned rollout_hash(unsigned long long seed, unsigned env, unsigned step) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * ((static_cast<unsigned long long>(env) << 20) ^ step ^ 0x5851F42D4C957F2Dull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return static_cast<unsigned>(z >> 32);
}

constexpr int kNumStats = 15;
// index of each statistic in RolloutArgs::stats (weather2alert_b200/distributed.py: STAT_NAMES)
enum StatIndex {
  kStEpisodes = 0, kStReturnSum, kStReturnSq, kStSteps, kStAlerts, kStRefused,
  kStAlertDaySum, kStAlertDaySq, kStStreaks, kStStreakSum, kStStreakSq,
  kStA50Sum, kStA80Sum, kStA100Sum, kStACount
};

// POLICY = false: open-loop actions (bank or Bernoulli hash).  POLICY = true: the action of day d is
// drawn by a logistic policy from the observation the agent holds before acting on day d -- the row of
// day d-1 (day 0 on days 0 and 1: the observation lags a day, quirk Q1) with the agent slots as the
// step of day d-1 returned them; only envs at the start of a summer (t = 0) take part.
// TAB = the table form (the whole-summer tables of w2a_precompute_logits exist): an env on its own county's column
// (canonical bit) reads (z_b, z_e) of its 8 days as ONE 128-byte run of zt_all -- 16 bytes per env-step instead of a
// 256-byte day-row and a 24-term dot product -- and the policy's table logit comes from pt[row].  The table kernel
// leaves non-canonical envs alone (RolloutArgs::only = 1); a second, direct launch (only = 2) takes exactly those.
template <bool POLICY, bool TAB>
__global__ void __launch_bounds__(128, TAB ? (POLICY ? 4 : 5) : (POLICY ? 3 : 4)) rollout_kernel(const RolloutArgs p) {
  __shared__ double s_stats[kNumStats];
  __shared__ unsigned char s_days[4][4][32];  // [warp][group][chunk of 8 days]: issued-alert bits of the episode so far (T <= 255)
  if (threadIdx.x < kNumStats) s_stats[threadIdx.x] = 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int q = lane & 7;
  const int g = lane >> 3;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const int n_quads = (p.n + 3) >> 2;  // 4 envs per warp iteration
  const double2* const x_lane = reinterpret_cast<const double2*>(p.tb.x) + q;
  unsigned char* const my_days = s_days[threadIdx.x >> 5][g];
  double acc[kNumStats];
#pragma unroll
  for (int i = 0; i < kNumStats; ++i) acc[i] = 0.0;

  for (int wq = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; wq < n_quads; wq += warps_total) {
    const int e = wq * 4 + g;
    const bool valid = e < p.n;
    const int ec = valid ? e : p.n - 1;
    const uint2 sv = __ldg(p.st_static + ec);
    const uint2 mv = ld_stream(p.st_mut + ec);  // written later in this launch: not through the read-only path
    const int n_days = sv.x >> 24;
    const unsigned row0 = sv.x & 0xffffffu;
    int rem = static_cast<int>(mv.x);
    int t = mv.y & 0xff;
    int streak = (mv.y >> 8) & 0xff;
    unsigned win = (mv.y >> 16) & 0x3fffu;
    bool done = ((mv.y >> 30) & 1u) != 0 || !valid;
    // the alert-timing quantiles need the whole summer: only for episodes that start in this call
    const bool whole_episode = t == 0;
    const bool canon = (mv.y >> 31) != 0;  // the env's column is its own county: its logits are tabulated
    if ((p.only == 1 && !canon) || (p.only == 2 && canon)) done = true;  // the other launch's env
    const double2* const zt_env =
        TAB ? p.tb.zt_all + (static_cast<size_t>(sv.y & kRecMask) * p.tb.Y + (sv.y >> 24)) * p.tb.T : nullptr;

    // coefficient record -> float64 registers, once per episode
    const float4* wr = reinterpret_cast<const float4*>(p.tb.w + static_cast<size_t>(sv.y & kRecMask) * kWStride);
    const float4 o0 = __ldg(wr + 14), o1 = __ldg(wr + 15);  // (b,e) of lag1, streak | remaining, bias
    // this lane's 8 table coefficients in float64 (direct form only)
    double wab0 = 0.0, wae0 = 0.0, wab1 = 0.0, wae1 = 0.0, wbb0 = 0.0, wbe0 = 0.0, wbb1 = 0.0, wbe1 = 0.0;
    if (!TAB) {
      const float4 wa4 = __ldg(wr + q), wb4 = __ldg(wr + 8 + q);
      wab0 = wa4.x; wae0 = wa4.y; wab1 = wa4.z; wae1 = wa4.w;
      wbb0 = wb4.x; wbe0 = wb4.y; wbb1 = wb4.z; wbe1 = wb4.w;
    }
    const double c_lag_b = o0.x, c_lag_e = o0.y, c_str_b = o0.z, c_str_e = o0.w;
    const double c_rem_b = o1.x, c_rem_e = o1.y, c_bias_b = o1.z, c_bias_e = o1.w;

    // ---- closed-loop policy: this lane's four table weights, the agent weights, what the agent last saw ----
    double th_a0 = 0.0, th_a1 = 0.0, th_b0 = 0.0, th_b1 = 0.0, th_lag = 0.0, th_streak = 0.0, th_rem = 0.0,
           th_2wks = 0.0, th_bias = 0.0;
    double pt_prev = 0.0;                  // policy's table logit of the row the agent saw last
    int lag_seen = 0, streak_seen = 0, cnt_seen = 0;  // agent slots of that observation (remaining budget = rem)
    if (POLICY) {
      // record position -> observation slot: 0..22 -> same, 23 -> 26 (issued_in_advance), 24..31 -> none
      const int pb0 = 16 + 2 * q, pb1 = 17 + 2 * q;
      th_a0 = __ldg(p.policy + 2 * q);
      th_a1 = __ldg(p.policy + 2 * q + 1);
      th_b0 = pb0 < 24 ? __ldg(p.policy + slot_of_feature(pb0)) : 0.0;
      th_b1 = pb1 < 24 ? __ldg(p.policy + slot_of_feature(pb1)) : 0.0;
      th_lag = __ldg(p.policy + 23);
      th_streak = __ldg(p.policy + 24);
      th_rem = __ldg(p.policy + 25);
      th_2wks = __ldg(p.policy + 28);
      th_bias = __ldg(p.policy + kObsDim);
      if (t != 0) done = true;  // mid-summer envs are left alone: their last observation is not in the state
    }

    double ret = 0.0;
    int k = 0;  // steps taken by this env in this call
    // table form: what lane q needs of a chunk -- (z_b, z_e) of day q, the policy's table logit, the action byte -- is
    // requested one chunk ahead, so that a chunk never starts by waiting for memory (the env's 8 days are ONE contiguous
    // 128-byte run of zt_all; the chunk's day range follows from t and k alone)
    double2 z_nxt = make_double2(0.0, 0.0);
    double pt_nxt = 0.0;
    int a_nxt = 0;
    auto prefetch_chunk = [&](int t0, int k0, bool alive) {
      if constexpr (TAB) {
        const int nv0 = alive && k0 < p.steps ? min(8, min(p.steps - k0, n_days - t0)) : 0;
        const int dq = min(t0 + q, n_days - 1);
        z_nxt = __ldg(zt_env + dq);
        if (POLICY) pt_nxt = __ldg(p.tb.pt + row0 + dq);
        if (!POLICY && p.actions) a_nxt = q < nv0 ? __ldg(p.actions + static_cast<size_t>(k0 + q) * p.n + e) : 0;
      }
    };
    prefetch_chunk(t, 0, !done);
    // the 4 envs of a warp may need different numbers of chunks: loop while any is alive
    while (__any_sync(kFull, !done && k < p.steps)) {
      const bool live = !done && k < p.steps;
      // days this env executes in this chunk: t .. t+nv-1, the last one possibly terminal
      const int nv = live ? min(8, min(p.steps - k, n_days - t)) : 0;
      const double2 z_tab = z_nxt;
      const double pt_tab = pt_nxt;
      const int a_tab = a_nxt;
      prefetch_chunk(t + nv, k + nv, live && (t + nv) < n_days);  // the next chunk's, in flight during this one
      // ---- table part of the logits of the chunk's days ----------------------------------
      double pbv[TAB ? 1 : 8], pev[TAB ? 1 : 8], ptv[POLICY && !TAB ? 8 : 1];
      if constexpr (!TAB) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const unsigned row = row0 + static_cast<unsigned>(min(t + j, n_days - 1));
          const double2* xr = x_lane + static_cast<size_t>(row) * (kXStride / 2);
          const double2 xa = __ldg(xr), xb = __ldg(xr + 8);
          pbv[j] = fma(xa.x, wab0, xa.y * wab1) + fma(xb.x, wbb0, xb.y * wbb1);
          pev[j] = fma(xa.x, wae0, xa.y * wae1) + fma(xb.x, wbe0, xb.y * wbe1);
          if (POLICY && !TAB) ptv[j] = fma(xa.x, th_a0, xa.y * th_a1) + fma(xb.x, th_b0, xb.y * th_b1);
        }
      }
      unsigned actm = 0, refm = 0;
      if (!POLICY) {
        // ---- this chunk's 8 action bits: lane q fetches day q's ------------------------------
        int a_mine = 0;
        if (q < nv) {
          if (p.actions) a_mine = TAB ? (a_tab != 0) : (__ldg(p.actions + static_cast<size_t>(k + q) * p.n + e) != 0);
          else a_mine = rollout_hash(p.seed, static_cast<unsigned>(e), static_cast<unsigned>(k + q)) < p.threshold;
        }
        const unsigned am = (__ballot_sync(kFull, a_mine) >> (8 * g)) & 0xffu;
        // ---- integer recurrence of the chunk (env.py:242-250) in closed form ----------------------
        // `at_budget = (remaining == 0)` refuses every alert once a POSITIVE budget has run out and never fires for a
        // negative one: the accepted alerts are the first `rem` requested ones (all of them when rem < 0 or enough is left)
        unsigned over = 0u;
        if (rem >= 0 && rem < __popc(am)) {
          over = am;
          for (int i = 0; i < rem; ++i) over &= over - 1u;  // drop the `rem` earliest requests: what is left is refused
        }
        actm = am & ~over;
        refm = over;
      } else {
        // ---- the policy's table logits of the chunk's rows, one per lane, then known to every lane ----
        double t2[2] = {pt_tab, 0.0};
        if constexpr (!TAB) {
          double t8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) t8[j] = ptv[j];
          fold_level<8>(t8, (q & 4) != 0, 4);
          double t4[4] = {t8[0], t8[1], t8[2], t8[3]};
          fold_level<4>(t4, (q & 2) != 0, 2);
          t2[0] = t4[0];
          t2[1] = t4[1];
          fold_level<2>(t2, (q & 1) != 0, 1);
        }
        double ptall[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) ptall[j] = __shfl_sync(kFull, t2[0], (lane & 24) + j);
        if (t == 0) pt_prev = ptall[0];  // the reset observation shows day 0
        // ---- decisions and integer recurrence, day after day (every lane replays them) ----------------
        int r = rem, s_cur = streak;
        unsigned w14 = win;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          double logit = (j == 0 ? pt_prev : ptall[j - 1]) + th_bias;
          logit = fma(static_cast<double>(lag_seen), th_lag, logit);
          logit = fma(static_cast<double>(streak_seen), th_streak, logit);
          logit = fma(static_cast<double>(r), th_rem, logit);
          logit = fma(static_cast<double>(cnt_seen), th_2wks, logit);
          unsigned a;
          if (p.stochastic) {
            const double thr = expit_fast(logit) * 4294967296.0;
            const unsigned thr_u = thr >= 4294967295.0 ? 0xffffffffu : static_cast<unsigned>(thr);
            a = rollout_hash(p.seed, static_cast<unsigned>(e), static_cast<unsigned>(k + j)) < thr_u ? 1u : 0u;
          } else {
            a = logit > 0.0 ? 1u : 0u;
          }
          if (j < nv) {
            const unsigned atb = (r == 0) ? 1u : 0u;
            const unsigned act = a & ~atb;
            r -= static_cast<int>(act);
            actm |= act << j;
            refm |= (a & atb) << j;
            // what the step of this day hands back (env.py:190-193): lag (0 on day 0), the streak before
            // today's action, the 14-day count with it
            lag_seen = (t + j) > 0 ? static_cast<int>(act) : 0;
            streak_seen = s_cur;
            s_cur = act ? s_cur + 1 : 0;
            w14 = ((w14 << 1) | act) & 0x3fffu;
            cnt_seen = __popc(w14);
          }
        }
        pt_prev = ptall[7];
      }
      // ---- transpose: lane q gets the table logits of day q ---------------------------------
      double b2[2] = {0.0, 0.0}, e2[2] = {0.0, 0.0};
      if constexpr (!TAB) {
        fold_level<8>(pbv, (q & 4) != 0, 4);
        fold_level<8>(pev, (q & 4) != 0, 4);
        double b4[4] = {pbv[0], pbv[1], pbv[2], pbv[3]}, e4[4] = {pev[0], pev[1], pev[2], pev[3]};
        fold_level<4>(b4, (q & 2) != 0, 2);
        fold_level<4>(e4, (q & 2) != 0, 2);
        b2[0] = b4[0]; b2[1] = b4[1];
        e2[0] = e4[0]; e2[1] = e4[1];
        fold_level<2>(b2, (q & 1) != 0, 1);
        fold_level<2>(e2, (q & 1) != 0, 1);
      }
      // ---- lane q evaluates day q (env.py:190-193, 206-221; quirk Q1) --------------------------
      double r_mine = 0.0;
      if (q < nv) {
        const unsigned below = actm & ((1u << q) - 1u);
        const int act = (actm >> q) & 1;
        const int rem_after = rem - __popc(actm & ((2u << q) - 1u));
        const unsigned zeros = ~below & ((1u << q) - 1u);
        const int streak_q = zeros == 0u ? streak + q : q - 1 - (31 - __clz(zeros));
        const int day = t + q;
        const int lag = day > 0 ? act : 0;
        // the tabulated pair already holds bias and gate; the folded dot products hold the gate only
        double zb = TAB ? z_tab.x : b2[0] + c_bias_b, ze = TAB ? z_tab.y : e2[0] + c_bias_e;
        zb = fma(static_cast<double>(lag), c_lag_b, zb);
        ze = fma(static_cast<double>(lag), c_lag_e, ze);
        zb = fma(static_cast<double>(streak_q), c_str_b, zb);
        ze = fma(static_cast<double>(streak_q), c_str_e, ze);
        zb = fma(static_cast<double>(rem_after), c_rem_b, zb);
        ze = fma(static_cast<double>(rem_after), c_rem_e, ze);
        const double baseline = expit_fast(zb);
        const double eff = act ? expit_fast(ze) : 0.0;
        r_mine = kRewardScale * baseline * (1.0 - eff);
        // timing statistics of this day: when alerts are issued, how long streaks of issued alerts last
        if (act) {
          acc[kStAlertDaySum] += day;
          acc[kStAlertDaySq] += static_cast<double>(day) * day;
        } else if (streak_q > 0) {  // a streak of streak_q issued alerts ended yesterday
          acc[kStStreaks] += 1.0;
          acc[kStStreakSum] += streak_q;
          acc[kStStreakSq] += static_cast<double>(streak_q) * streak_q;
        }
      }
      r_mine += __shfl_xor_sync(kFull, r_mine, 4);
      r_mine += __shfl_xor_sync(kFull, r_mine, 2);
      r_mine += __shfl_xor_sync(kFull, r_mine, 1);
      ret += r_mine;
      // ---- carry the state over the chunk (env.py:256-260) -------------------------------------
      if (nv > 0) {
        const unsigned vmask = (1u << nv) - 1u;  // nv <= 8
        const bool terminal = (t + nv) >= n_days;  // the chunk's last day is the episode's last
        if (whole_episode && q == 0) my_days[t >> 3] = static_cast<unsigned char>(actm & vmask);  // t is a multiple of 8 here
        // the streak is not updated on the terminal day
        const int ns = terminal ? nv - 1 : nv;
        const unsigned sm = (1u << ns) - 1u;
        const unsigned zeros = ~actm & sm;
        streak = zeros == 0u ? streak + ns : ns - 1 - (31 - __clz(zeros));
        rem -= __popc(actm & vmask);
        // the window keeps the most recent action in bit 0, the chunk mask has day j in bit j
        win = ((win << nv) | (__brev(actm & vmask) >> (32 - nv))) & 0x3fffu;
        t = terminal ? n_days - 1 : t + nv;
        k += nv;
        if (q == 0) {
          acc[kStSteps] += nv;
          acc[kStAlerts] += __popc(actm & vmask);
          acc[kStRefused] += __popc(refm & vmask);
        }
        if (terminal) done = true;
      }
    }
    if (valid && q == 0 && k > 0) {
      const double total = p.ep_return[e] + ret;
      p.ep_return[e] = total;
      p.st_mut[e] = make_uint2(static_cast<unsigned>(rem),
                               static_cast<unsigned>(t) | (static_cast<unsigned>(streak) << 8) | (win << 16) |
                                   ((done ? 1u : 0u) << 30) | (mv.y & 0x80000000u));
      if (done) {
        acc[kStEpisodes] += 1.0;
        acc[kStReturnSum] += total;
        acc[kStReturnSq] += total * total;
        if (whole_episode) {
          // day by which 50 % / 80 % / 100 % of the summer's alerts had been issued (callbacks.py:47-57)
          const int n_chunks = (n_days + 7) >> 3;
          int issued = 0;
          for (int c = 0; c < n_chunks; ++c) issued += __popc(static_cast<unsigned>(my_days[c]));
          if (issued > 0) {
            int cum = 0, d50 = -1, d80 = -1, d100 = -1;
            for (int d = 0; d < n_days && d100 < 0; ++d) {
              cum += (my_days[d >> 3] >> (d & 7)) & 1;
              if (d50 < 0 && 2 * cum >= issued) d50 = d;
              if (d80 < 0 && 5 * cum >= 4 * issued) d80 = d;
              if (cum == issued) d100 = d;
            }
            acc[kStA50Sum] += d50;
            acc[kStA80Sum] += d80;
            acc[kStA100Sum] += d100;
            acc[kStACount] += 1.0;
          }
        }
      }
    }
    __syncwarp();
  }
  // ---- statistics: lanes -> warp (shuffles) -> block (shared atomics) -> one global atomic per block and field
  if (p.stats) {
#pragma unroll
    for (int i = 0; i < kNumStats; ++i) {
      double v = acc[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
      if (lane == 0 && v != 0.0) atomicAdd(&s_stats[i], v);
    }
    __syncthreads();
    if (threadIdx.x < kNumStats && s_stats[threadIdx.x] != 0.0) atomicAdd(p.stats + threadIdx.x, s_stats[threadIdx.x]);
  }
}

}  // namespace w2a
