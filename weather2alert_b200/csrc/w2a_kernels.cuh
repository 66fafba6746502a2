// w2a_kernels.cuh -- sm_100a kernels of the HeatAlertEnv hot path.
//
// What the kernels compute is the reference's HeatAlertEnv.step / reset
// (src/weather2alert/env.py:133-262 of NSAPH-Projects/weather2alert) for a whole batch
// of independent envs; how they compute it is laid out for B200:
//
//   * per-env state is two SoA planes of 8 bytes (static: episode row + coefficient
//     vector; mutable: remaining budget + packed day/streak/14-day window), so a warp
//     reads and writes it with fully coalesced 8-byte accesses;
//   * both tables are arrays of 256-byte, line-aligned records (one day-row of float64
//     features; one float32 coefficient vector), so the gather of one env is exactly
//     four 128-byte lines, each fetched by 8 lanes with one 128-bit load apiece;
//   * the step kernel gives every env 8 lanes for that gather, accumulates the two
//     28-term dot products in float64, folds the 8 partials through shared memory, and
//     stages the 29-slot observations of a warp's 32 envs in shared memory so that they
//     leave as one contiguous 3712-byte run of 128-bit streaming stores;
//   * tables are read through the read-only path and stay L2 resident; the per-env
//     streams (state, action, obs, reward, flags) use streaming (.cs) accesses so they
//     do not evict the tables.
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
// env.py:221 -- the reference divides by 152, not 153
constexpr double kRewardScale = -1000.0 / 152.0;

struct Tables {
  // [R, 32] float64: 0..23 = reward features read from the table (obs slots 0..22 and 26),
  // 24 = significance code (0.0 = none, k = label k), 25..31 = 0.0
  const double* x;
  const uint8_t* n_days;   // [S*Y]
  const int32_t* budget0;  // [S*Y]
  // [Sw*P, 64] float32: (baseline, effectiveness) pairs; 0..47 = the 24 table features,
  // 48..55 = 0, 56..63 = alert_lag1, alert_streak, remaining_budget, bias
  const float* w;
  int S, Y, T, P, Sw;
};

struct StepArgs {
  Tables tb;
  const uint2* st_static;  // x: row0 (24 bit) | n_days << 24 ; y: coefficient record index
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

// ---- helpers --------------------------------------------------------------------------
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
         (static_cast<unsigned>(o.done) << 30);
  return o;
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

    const double* __restrict__ xr = p.tb.x + static_cast<size_t>(tr.row) * kXStride;
    const float* __restrict__ wv = p.tb.w + static_cast<size_t>(sv.y) * kWStride;
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
    const double baseline = expit(zb);
    const double eff = (tr.act && heat_qi > 0.5) ? expit(ze) : 0.0;
    p.reward[e] = static_cast<float>(kRewardScale * baseline * (1.0 - eff));
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
// B) of the day-row and the matching (b,e) coefficient pairs.  Positions 24..31 carry no
// table feature: B of lanes 4..7 is (sig, 0) / zeros on the x side, zeros (lanes 4,5) or
// the non-table coefficients (lanes 6,7) on the w side, so their products vanish.  For
// the observation a lane owns these slots:
//   A.x -> 2q      A.y -> 2q+1
//   B.x -> 16+2q (q<4) | 27 significance (q=4) | 23 alert_lag1 (q=5) | 25 remaining_budget (q=6)
//   B.y -> 17+2q (q<3) | 26 (q=3) | 28 alert_2wks (q=4) | 24 alert_streak (q=5)
// =========================================================================================
struct LaneSlots {
  int bx, by;  // observation slot written from B.x / B.y, -1 = none
};
__device__ __forceinline__ LaneSlots lane_slots(int q) {
  LaneSlots s;
  s.bx = q < 4 ? 16 + 2 * q : (q == 4 ? 27 : (q == 5 ? 23 : (q == 6 ? 25 : -1)));
  s.by = q < 3 ? 17 + 2 * q : (q == 3 ? 26 : (q == 4 ? 28 : (q == 5 ? 24 : -1)));
  return s;
}

#ifndef W2A_COOP_WARPS
#define W2A_COOP_WARPS 8  // warps per CTA
#endif
#ifndef W2A_COOP_MINBLOCKS
#define W2A_COOP_MINBLOCKS 3  // CTAs per SM the register allocation is sized for
#endif
constexpr int kCoopWarps = W2A_COOP_WARPS;

// per-warp shared memory of the COOP kernel
struct __align__(16) WarpScratch {
  double2 part[8][32];      // [round][lane ^ round]: (pb, pe) partial dot products
  float4 other[4][2][8];    // [group][lane-6][round ^ (2*group + lane-6)]: raw non-table coefficients for the env's owner
  float obs[32 * kObsDim];  // the warp's 32 observation rows, same layout as in global memory
};
constexpr int kCoopSmemBytes = kCoopWarps * static_cast<int>(sizeof(WarpScratch));

struct RoundLoads {
  double2 xa, xb;
  float4 wa, wb;
};

__device__ __forceinline__ RoundLoads issue_round(const double2* x_lane, const float4* w_lane, unsigned row,
                                                  unsigned cv) {
  RoundLoads r;
  const double2* xr = x_lane + static_cast<size_t>(row) * (kXStride / 2);
  const float4* wr = w_lane + static_cast<size_t>(cv) * (kWStride / 4);
  r.xa = __ldg(xr);
  r.xb = __ldg(xr + 8);
  r.wa = __ldg(wr);
  r.wb = __ldg(wr + 8);
  return r;
}

// =========================================================================================
// step, variant COOP: a warp owns 32 consecutive envs.
//
//   phase A  one thread per env: coalesced state/action traffic and the integer transition.
//   phase B  eight rounds, four envs per round, 8 lanes per env: four 128-bit loads per
//            lane fetch the env's day-row and coefficient record (the next round's loads
//            are issued before this round's math); the lane's float64 partial dot products
//            and its observation slots go to shared memory.  No branches: the heat_qi gate
//            is folded into the effectiveness logit as -inf.
//   phase C  one thread per env: it sums the 8 partials of its env (XOR-swizzled, conflict
//            free), adds the three agent terms and the bias, evaluates the two float64
//            sigmoids and writes reward and flags; then the warp streams its 32 staged
//            observation rows out as one contiguous run.
// =========================================================================================
template <int ADT>
__global__ void __launch_bounds__(32 * kCoopWarps, W2A_COOP_MINBLOCKS) step_coop_kernel(const StepArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WarpScratch& ws = reinterpret_cast<WarpScratch*>(smem_raw)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int q = lane & 7;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const int n_groups = (p.n + 31) >> 5;
  const double2* const x_lane = reinterpret_cast<const double2*>(p.tb.x) + q;
  const float4* const w_lane = reinterpret_cast<const float4*>(p.tb.w) + q;
  const LaneSlots ls = lane_slots(q);
  float* const stage_lane = ws.obs + (lane & 24) * kObsDim;  // rows of this lane's 8-env group
  double2* const part0 = &ws.part[0][lane];
  // agent counters shown in the observation: B.y of lane 5 = streak, of lane 4 = alert_2wks
  const int by_shift = q == 5 ? 8 : 16;
  const int by_mask = q == 5 ? 0xff : 0xf;
  const int oth_swz = 2 * (lane >> 3) + (q & 1);  // bank swizzle of `other` (lanes 6,7 write it)

  for (int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; wg < n_groups; wg += warps_total) {
    const int base = wg << 5;
    const int e = base + lane;
    const bool valid = e < p.n;
    const int ec = valid ? e : p.n - 1;

    // ---- phase A ---------------------------------------------------------------------------
    const uint2 sv = ld_stream(p.st_static + ec);
    uint2 mv = ld_stream(p.st_mut + ec);
    const int a = load_action<ADT>(p.action, ec);
    const Transition tr = advance(sv, mv, a);
    if (valid) st_stream(p.st_mut + e, mv);
    const unsigned cv = sv.y;
    const bool write_obs = valid && !tr.done;  // on and after the last day the previous obs stands (Q2)
    const int pk = tr.lag | (tr.streak << 8) | (tr.cnt14 << 16);
    const unsigned wmask = __ballot_sync(kFull, write_obs);

    // ---- phase B ---------------------------------------------------------------------------
    unsigned rowj = __shfl_sync(kFull, tr.row, 0, 8);
    unsigned cvj = __shfl_sync(kFull, cv, 0, 8);
    RoundLoads cur = issue_round(x_lane, w_lane, rowj, cvj);
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int pkj = __shfl_sync(kFull, pk, it, 8);
      const int remj = __shfl_sync(kFull, tr.rem, it, 8);
      RoundLoads nxt;
      if (it < 7) {
        rowj = __shfl_sync(kFull, tr.row, it + 1, 8);
        cvj = __shfl_sync(kFull, cv, it + 1, 8);
        nxt = issue_round(x_lane, w_lane, rowj, cvj);
      }
      double pb = cur.xa.x * static_cast<double>(cur.wa.x);
      double pe = cur.xa.x * static_cast<double>(cur.wa.y);
      pb = fma(cur.xa.y, static_cast<double>(cur.wa.z), pb);
      pe = fma(cur.xa.y, static_cast<double>(cur.wa.w), pe);
      pb = fma(cur.xb.x, static_cast<double>(cur.wb.x), pb);
      pe = fma(cur.xb.x, static_cast<double>(cur.wb.y), pe);
      pb = fma(cur.xb.y, static_cast<double>(cur.wb.z), pb);
      pe = fma(cur.xb.y, static_cast<double>(cur.wb.w), pe);
      // effectiveness counts only when heat_qi > 0.5 (env.py:218; Q4): heat_qi is lane 0's
      // A.x; -inf makes sigmoid(z_e) exactly 0
      if (q == 0 && !(cur.xa.x > 0.5)) pe = -CUDART_INF;
      part0[it * 32 + ((lane ^ it) - lane)] = make_double2(pb, pe);
      if (q >= 6) ws.other[lane >> 3][q - 6][it ^ oth_swz] = cur.wb;

      // observation of day t (Q2): float32 of the table row + the four agent slots
      float* const st = stage_lane + it * kObsDim;
      float vbx = static_cast<float>(cur.xb.x), vby = static_cast<float>(cur.xb.y);
      if (q == 4) vbx = sig_to_obs(vbx);
      if (q == 5) vbx = static_cast<float>(pkj & 1);
      if (q == 6) vbx = static_cast<float>(remj);
      if (q == 4 || q == 5) vby = static_cast<float>((pkj >> by_shift) & by_mask);
      st[2 * q] = static_cast<float>(cur.xa.x);
      st[2 * q + 1] = static_cast<float>(cur.xa.y);
      if (ls.bx >= 0) st[ls.bx] = vbx;
      if (ls.by >= 0) st[ls.by] = vby;
      if (it < 7) cur = nxt;
    }
    __syncwarp();

    // ---- phase C ---------------------------------------------------------------------------
    // env (lane) was round q of its 8-lane group; its partials sit in part[q][(g*8+j)^q]
    {
      const int g = lane >> 3;
      const float4 o0 = ws.other[g][0][q ^ (2 * g)];      // (b,e) of alert_lag1, alert_streak
      const float4 o1 = ws.other[g][1][q ^ (2 * g + 1)];  // (b,e) of remaining_budget, bias
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
        const double2 v = ws.part[q][(lane & 24) | (j ^ q)];
        zb += v.x;
        ze += v.y;
      }
      if (valid) {
        // reward = -1000/152 * sigmoid(z_b) * (1 - sigmoid(z_e) * [heat_qi > 0.5] * action)  (env.py:211-221)
        const double baseline = expit(zb);
        const double eff = tr.act ? expit(ze) : 0.0;
        __stcs(p.reward + e, static_cast<float>(kRewardScale * baseline * (1.0 - eff)));
        p.terminated[e] = static_cast<uint8_t>(tr.done);
        if (p.remaining) __stcs(p.remaining + e, tr.rem);
        if (p.at_budget) p.at_budget[e] = static_cast<uint8_t>(tr.at_budget);
      }
    }
    // observation rows: 32 x 116 B contiguous in global memory (base is a multiple of 32)
    float* const ob = p.obs + static_cast<size_t>(base) * kObsDim;
    if (wmask == kFull && ((reinterpret_cast<uintptr_t>(p.obs) & 15) == 0)) {
      const float4* src = reinterpret_cast<const float4*>(ws.obs);
      float4* dst = reinterpret_cast<float4*>(ob);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = k * 32 + lane;
        if (i < 32 * kObsDim / 4) __stcs(dst + i, src[i]);
      }
    } else if (wmask != 0u) {  // ragged tail or some envs already on their last day
      for (int i = lane; i < 32 * kObsDim; i += 32) {
        const int env = (i * 565) >> 14;  // i / 29 for i < 928
        if ((wmask >> env) & 1u) st_stream(ob + i, ws.obs[i]);
      }
    }
    __syncwarp();
  }
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
        const double2* xr = x_lane + static_cast<size_t>(rowj) * (kXStride / 2);
        const double2 xa = __ldg(xr), xb = __ldg(xr + 8);
        float vbx = static_cast<float>(xb.x), vby = static_cast<float>(xb.y);
        if (q == 4) vbx = sig_to_obs(vbx);
        if (q == 5) vbx = 0.f;                        // alert_lag1 (t == 0)
        if (q == 6) vbx = static_cast<float>(budj);   // remaining_budget = budget
        if (q == 4 || q == 5) vby = 0.f;              // alert_2wks, alert_streak
        float* ob = p.obs + static_cast<size_t>(base + (lane & 24) + it) * kObsDim;
        st_stream(ob + 2 * q, static_cast<float>(xa.x));
        st_stream(ob + 2 * q + 1, static_cast<float>(xa.y));
        if (ls.bx >= 0) st_stream(ob + ls.bx, vbx);
        if (ls.by >= 0) st_stream(ob + ls.by, vby);
      }
    }
  }
}

}  // namespace w2a
