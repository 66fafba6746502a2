/*
 * w2a.h -- C ABI of the B200-native HeatAlertEnv reset/step hot path.
 *
 * The reference (NSAPH-Projects/weather2alert) has no FFI of its own: the path sits
 * behind a Gymnasium `Env` written in Python.  Every entry point below therefore cites
 * the Python method it replaces (paths relative to the reference checkout); the Python
 * classes in `weather2alert_b200/` bind these symbols with ctypes and mirror the
 * reference's signatures, and INTEGRATION.md shows the stub a reference maintainer
 * would add.
 *
 * Conventions
 *   - every function returns 0 on success or a negative W2A_E_* code; the message for
 *     the last failure on a context is returned by w2a_last_error();
 *   - the caller owns every buffer; the library never allocates outputs nor frees
 *     inputs (the *_host entry points own only their private staging buffers);
 *   - "dev" pointers are device pointers on the context's GPU; kernels are enqueued on
 *     the caller-supplied `stream` (a cudaStream_t passed as void*, NULL = default
 *     stream) and the call returns without synchronising;
 *   - a context is not thread-safe; no C++ exception crosses this boundary.
 */
#ifndef W2A_H_
#define W2A_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define W2A_VERSION 210 /* 0.2.0: + w2a_precompute_logits (whole-summer tables; W2A_STEP_TABLE no longer builds per step),
                           w2a_engineer_features, w2a_policy_mlp; 0.2.1: + w2a_step_mlp */

#define W2A_OK 0
#define W2A_E_INVALID (-1)  /* bad argument (null pointer, size, dtype, variant) */
#define W2A_E_CUDA (-2)     /* CUDA runtime error; text in w2a_last_error */
#define W2A_E_NOTREADY (-3) /* tables / host env set not installed yet */
#define W2A_E_NOMEM (-4)

/* Observation layout: 29 float32 slots per env, row-major [n, 29]; slot order is
 * `info["feature_names"]` of the reference (src/weather2alert/env.py:182-183). */
#define W2A_OBS_DIM 29
/* Device table records (see DESIGN.md "Data layout in HBM"): both are 256 bytes, and the
 * tables must be 128-byte aligned so that a record is exactly two cache lines. */
#define W2A_X_STRIDE 32 /* float64 per day-row: 0..23 table reward features, 24 significance code, 25 gate, 26..31 zero */
#define W2A_W_STRIDE 64 /* float32 per (county, posterior sample): (baseline, effectiveness) pairs */

/* OR into `variant`: the env slots are in table order (e.g. sorted by county, sample, year), so that neighbouring envs
 * share table lines -- W2A_STEP_TABLE then keeps two rounds of row gathers in flight at every size (a tuning hint:
 * results never depend on it) */
#define W2A_STEP_LOCAL 0x100

/* action_dtype */
#define W2A_ACT_U8 0
#define W2A_ACT_I32 1
#define W2A_ACT_I64 2

/* kernel variant for w2a_step */
#define W2A_STEP_AUTO 0   /* = COOP (it is the faster of the two at every batch size, a single env included) */
#define W2A_STEP_THREAD 1 /* one thread per env: the straightforward kernel, kept as the in-tree cross-check */
#define W2A_STEP_COOP 2   /* 8 lanes per env: whole-line gathers, float64 fold in shared memory, TMA obs store;
                             small batches spread their warps over the SMs with every gather in flight at once */
#define W2A_STEP_TABLE 3  /* one 128-byte record of tabulated logits + one float32 observation row per env, from the
                             whole-summer tables of w2a_precompute_logits; every env looks its own day up (no lock
                             step needed); for env sets that step on their own county's coefficients -- envs the
                             table does not cover (similar-county columns, quirk Q6) are computed directly, so
                             results never depend on the variant */

/* per-env budget value meaning "use the episode's historical budget"
 * (env.py:167-170: `self.ep["remaining_budget"].iloc[0]`). */
#define W2A_BUDGET_DEFAULT INT32_MIN
#define W2A_NUM_STATS 15

typedef struct w2a_ctx w2a_ctx;

int w2a_version(void);
/* Message of the last failure on `ctx` (or of the last failed w2a_create when ctx is
 * NULL).  Never NULL. */
const char* w2a_last_error(const w2a_ctx* ctx);

/* One context per device.  Replaces the table-holding half of
 * HeatAlertEnv.__init__ (env.py:20-105). */
int w2a_create(int device, w2a_ctx** out);
int w2a_destroy(w2a_ctx* ctx);

/* Install the HBM-resident feature store (replaces `self.merged`, env.py:49-56).
 *   x       dev float64 [S*Y*T, 32]  one record per (county, year, day): positions 0..23 are
 *                                    the reward features read from the table (observation slots
 *                                    0..22 and 26, float64 as in the parquet files), position 24
 *                                    the significance code (0.0 = none, shown as NaN; k = label
 *                                    k), position 25 the heat_qi gate (0.0 if heat_qi > 0.5 else
 *                                    -1e300; env.py:218), positions 26..31 must be 0.0
 *   n_days  dev uint8   [S*Y]        days in episode (county, year); 0 = no such episode
 *   budget0 dev int32   [S*Y]        historical `remaining_budget` of day 0 (env.py:169)
 * The tables are referenced, not copied, and must outlive their use.  S*Y*T < 2^24, Y <= 255, T <= 255. */
int w2a_set_features(w2a_ctx* ctx, const double* x, const uint8_t* n_days, const int32_t* budget0,
                     int S, int Y, int T);

/* Install the posterior coefficient draws (replaces `self.baseline_coefs` /
 * `self.effectiveness_coefs`, env.py:77-82).
 *   w  dev float32 [S_w*P, 64]; record (col*P + p) holds the coefficients of county column
 *      `col`, posterior sample `p` as (baseline, effectiveness) pairs: floats 0..47 the 24
 *      table features in record order, 48..55 must be 0 except float 51 = 1.0 (effectiveness
 *      coefficient of the gate column), 56..63 alert_lag1, alert_streak, remaining_budget, bias.
 *      S_w * P < 2^24. */
int w2a_set_coefs(w2a_ctx* ctx, const float* w, int P, int S_w);

/* Optional: tell the library which feature-store county each coefficient column belongs to
 * (`fips_list[col]` -> index into the feature store, -1 = not in the store; the reference
 * keeps both keyed by FIPS string, env.py:75,121,127).  Call after w2a_set_features /
 * w2a_set_coefs and before w2a_reset (the reset marks the envs whose column is their own
 * county).  `col_to_loc` is dev int32 [S_w], referenced, not copied (the call reads it once,
 * synchronously, for the host-side checks of the *_host entry points). */
int w2a_set_county_map(w2a_ctx* ctx, const int32_t* col_to_loc);

/* Hoist the agent-independent part of the reward out of `step` (replaces the per-key loop of
 * HeatAlertEnv._get_reward, env.py:206-217, for the 24 table features, the bias and the heat_qi
 * gate): for every day of the summer, every (coefficient column, posterior sample) and every year
 * of the column's county, z_b / z_e in float64 (a float64 GEMM per county and day on the FP64 tensor
 * path), plus the float32 observation row of every day-row (the row copy of _get_obs, env.py:187).
 * Two forms are kept: one 32-byte sector per (day, column, sample, year) with the six agent
 * coefficients and (z_b, z_e) rounded to float32 -- what the step kernel gathers, one L2 request per
 * env -- and the float64 pairs in episode-major order, which the step kernel falls back to wherever
 * the float32 rounding could reach the reward (|z| >= 32 that the agent terms pull back out of
 * saturation) and which w2a_rollout / w2a_rollout_policy stream instead of the day-rows.
 * The context allocates T*S_w*P*Y*(32 + 16) + S*Y*T*136 bytes (w2a_precompute_bytes;
 * 6.2 GB for 746 counties x 11 years x 153 days x 100 samples) and enqueues two kernels on `stream`;
 * the tables are valid for work enqueued on that stream afterwards (or after a synchronisation) and
 * are read-only from then on, so W2A_STEP_TABLE launches of one context may run on several streams.
 * Enables W2A_STEP_TABLE.  Needs w2a_set_county_map; any later w2a_set_* call drops the tables.
 * Returns W2A_E_NOMEM (nothing allocated) when the device does not have the room. */
int w2a_precompute_logits(w2a_ctx* ctx, void* stream);
size_t w2a_precompute_bytes(const w2a_ctx* ctx);

/* Bytes of the opaque per-env state for `n` envs (16 bytes per env, two SoA planes). */
size_t w2a_state_bytes(int64_t n);

/* Reset `n` envs from host-sampled index streams (replaces HeatAlertEnv.reset +
 * _get_episode, env.py:107-184, after the RNG draws).  All arrays are dev, length n.
 *   loc       county index into the feature store (episode weather; quirk Q7)
 *   year      year index into the feature store
 *   coef_col  column into the coefficient tensors: fips_list index, or the index inside
 *             the similar-county list when augmenting (env.py:117,199; quirk Q6)
 *   coef_idx  posterior sample (env.py:160)
 *   budget    alert budget, W2A_BUDGET_DEFAULT per env (or NULL for all) = historical
 *   mask      NULL, or uint8 per env: only envs with mask != 0 are reset
 * Writes state, the first observation and (if non-NULL) remaining[]. */
int w2a_reset(w2a_ctx* ctx, void* stream, int64_t n, const int32_t* loc,
              const int32_t* year, const int32_t* coef_col, const int32_t* coef_idx,
              const int32_t* budget, const uint8_t* mask, void* state, float* obs,
              int32_t* remaining);

/* One env.step for `n` envs in one launch (replaces HeatAlertEnv.step + _get_reward +
 * _get_obs + _get_info, env.py:186-262).
 *   action      dev [n], dtype per action_dtype; non-zero = issue an alert
 *   state       in/out
 *   obs         in/out dev float32 [n,29]: rows of envs that are already on their last
 *               day are left untouched, exactly as the reference returns the previous
 *               observation object on and after the terminal step (env.py:257-262)
 *   reward      out dev float32 [n]
 *   terminated  out dev uint8 [n]
 *   remaining   out dev int32 [n] or NULL  (info["remaining_budget"], post-action)
 *   at_budget   out dev uint8 [n] or NULL  (info["at_budget"], pre-action) */
int w2a_step(w2a_ctx* ctx, void* stream, int64_t n, const void* action, int action_dtype,
             void* state, float* obs, float* reward, uint8_t* terminated,
             int32_t* remaining, uint8_t* at_budget, int variant);

/* The whole `while not done: env.step(a)` loop (env.py:273-276) of `n` envs in one launch:
 * up to `steps` steps per env, stopping at the env's terminal step; envs that are already
 * done are left alone.  Observations are not materialised.
 *   actions    dev uint8 [steps, n] row-major (row k = the k-th step of this call), or NULL:
 *              then action(env, k) = hash(seed, env, k) < p_alert * 2^32 (Bernoulli(p_alert),
 *              counter based -- see rollout_hash in w2a_kernels.cuh)
 *   state      in/out
 *   ep_return  dev float64 [n], += the float64 sum of this call's rewards of the env
 *   stats      dev float64 [W2A_NUM_STATS] or NULL, += the sufficient statistics of the reference's
 *              AlertLoggingCallback (callbacks.py:63-74), ready for one all-reduce across GPUs:
 *              0 episodes finished, 1-2 sum / sum of squares of their ep_return, 3 steps,
 *              4 alerts issued, 5 alerts refused at budget, 6-7 sum / sum of squares of the day
 *              of summer of every issued alert, 8-10 count / sum / sum of squares of the lengths
 *              of streaks of issued alerts that ended, 11-13 sums of the day by which 50 %, 80 %,
 *              100 % of a summer's alerts had been issued, 14 summers counted in 11-13 (those
 *              with at least one alert that ran from day 0 within this call)
 * With the tables of w2a_precompute_logits installed the call is two launches: envs on their own county's column
 * stream 16 bytes of tabulated float64 logits per env-step (an episode's days are one contiguous run), the others
 * (similar-county columns, quirk Q6) go through the day-rows as before; every env belongs to exactly one launch. */
int w2a_rollout(w2a_ctx* ctx, void* stream, int64_t n, int steps, const uint8_t* actions,
                uint64_t seed, float p_alert, void* state, double* ep_return, double* stats);

/* The same loop with the agent inside the launch (`a = policy(obs); env.step(a)`, env.py:273-276 with a
 * policy in place of `action_space.sample()`): a logistic policy on the 29-slot observation,
 *   logit = sum_s policy[s] * obs[s] (s != 27, the significance label) + policy[29],
 * deterministic (action = logit > 0) or stochastic (action = hash(seed, env, step) < sigmoid(logit) * 2^32,
 * the counter-based stream of w2a_rollout).  The observation is the one the reference agent holds before
 * acting: the first observation of the reset, then what each step returned (row of the day just stepped,
 * agent slots after the action; env.py:186-195, quirk Q1), in float64.  Only envs at the start of a summer
 * (fresh from w2a_reset) take part; others are left untouched.  With the tables installed the policy's logit over
 * the 24 table slots is tabulated per day-row first (one pass over the feature store into a context-owned buffer:
 * one w2a_rollout_policy at a time per context).
 *   policy   dev float64 [W2A_POLICY_LEN] */
#define W2A_POLICY_LEN 30
int w2a_rollout_policy(w2a_ctx* ctx, void* stream, int64_t n, int steps, const double* policy, int stochastic,
                       uint64_t seed, void* state, double* ep_return, double* stats);

/* A small MLP policy on the observation, for the step-by-step loop `a = policy(obs); env.step(a)` (env.py:273-276 with a
 * policy in place of `action_space.sample()`): one launch computes
 *   logit = w2 . act(W1 . nan_to_num(obs) + b1) + b2,   action = logit > 0
 * (or, stochastic, action = hash(seed, env, step) < sigmoid(logit) * 2^32 -- the counter-based stream of w2a_rollout)
 * for n envs.  Both layers run on the tensor cores with float16 inputs and float32 accumulation (obs, W1, b1, the hidden
 * activations and w2 are rounded to 11 significant bits, values beyond +-65504 clamped; tanh is the hardware
 * approximation, 2^-11 relative; evaluate the policy the same way when training on these rollouts).
 *   obs      dev float32 [n, 29]   (what w2a_step / w2a_reset wrote; NaN slots count as 0)
 *   w1       dev float32 [hidden, 29] row-major, b1 / w2 dev float32 [hidden]; hidden: multiple of 8, at most 64
 *   action   out dev uint8 [n];  logit  out dev float32 [n] or NULL */
#define W2A_ACT_RELU 0
#define W2A_ACT_TANH 1
int w2a_policy_mlp(w2a_ctx* ctx, void* stream, int64_t n, const float* obs, const float* w1, const float* b1,
                   const float* w2, float b2, int hidden, int activation, int stochastic, uint64_t seed, uint32_t step,
                   uint8_t* action, float* logit);

/* One iteration of that loop in ONE launch: the step of `action` (exactly w2a_step's outputs, `variant` as there) AND
 * the policy's choice for the NEXT step, evaluated on the fresh observations while they are still on chip:
 *   obs, reward, ... = step(action);   next_action = policy(obs)   [logit, ep_return += reward: optional]
 * The observation tensor is written once and not read back, which the two-launch form (w2a_step, w2a_policy_mlp) pays
 * for with 116 B per env through HBM.  With W2A_STEP_TABLE and the tables installed this is one kernel; any other
 * variant runs as the separate launches, with identical results.  Actions are uint8; `next_action` may not alias `action`
 * (alternate two buffers).  `step` numbers the policy's stochastic draw as in w2a_policy_mlp (the draw for the
 * action that step k + 1 consumes).
 *   ep_return  in/out dev float64 [n] or NULL */
int w2a_step_mlp(w2a_ctx* ctx, void* stream, int64_t n, const uint8_t* action, void* state, float* obs, float* reward,
                 uint8_t* terminated, int32_t* remaining, uint8_t* at_budget, int variant, const float* w1, const float* b1,
                 const float* w2, float b2, int hidden, int activation, int stochastic, uint64_t seed, uint32_t step,
                 uint8_t* next_action, float* logit, double* ep_return);

/* ---- host-buffer entry points: the call a CPU-side RL loop makes ------------------
 * The context owns a device-resident env set; every call copies its inputs from and its
 * outputs to HOST memory (pinned for full PCIe speed), pipelined in chunks over several
 * streams, and returns after the outputs have landed.  The step variant is chosen by the library: the
 * tabulated logits (W2A_STEP_TABLE) when the tables exist and the host index streams of the resets show
 * every env on its own county's column, else W2A_STEP_COOP.  Env sets of at most 1024 envs move their
 * actions and all step outputs through one pinned staging block (one copy each way), so the caller's
 * arrays may be ordinary pageable memory at no cost. */
int w2a_host_envs_create(w2a_ctx* ctx, int64_t n);
int w2a_host_envs_destroy(w2a_ctx* ctx);
int w2a_reset_host(w2a_ctx* ctx, const int32_t* loc, const int32_t* year,
                   const int32_t* coef_col, const int32_t* coef_idx,
                   const int32_t* budget, const uint8_t* mask, float* obs,
                   int32_t* remaining);
int w2a_step_host(w2a_ctx* ctx, const uint8_t* action, float* obs, float* reward,
                  uint8_t* terminated, int32_t* remaining, uint8_t* at_budget);
/* ---- feature engineering: the feature-store columns from raw daily records --------------
 * Replaces the pandas statements of the reference's ETL, data-processing/merge_state_actions.py:121-210: per-county
 * percentile rank of the heat index (:122), hinge and interaction terms (:125-133), 3/7-day rolling means and excess
 * heat (:136-147), day of summer (:151), alert history -- 14-day count, lag, streak (:154-172) --, weekend flag (:178),
 * remaining budget (:190-196) and the standardised cubic B-spline basis of the day of summer (:198-210).
 * Rows are sorted by (county, date) as at :23; all arrays are dev.
 *   county_start  int64 [n_counties + 1]   row offsets of the counties
 *   max_rows_per_county                    largest county, in rows (<= 6400: one county is staged in shared memory)
 *   date_days     int32 [n_rows]           days since 1970-01-01
 *   himax_c       float64 [n_rows]         daily maximum heat index in Celsius (`HImax_C`)
 *   alert         uint8 [n_rows]           an alert was in force that day (`issue.notnull()`, :114)
 *   issued_in_advance float64 [n_rows]     NaN = none (:175 fills 0)
 *   holiday       uint8 [n_rows]           host-supplied calendar flag (the reference asks the `holidays` package)
 *   feat          out float64 [n_rows, 27] the 27 numeric feature-store columns in merged-frame order (observation
 *                                          slots 0..26: `exogenous_states` then `endogenous_states_actions`, :228-272)
 *   year          out int32 [n_rows]       calendar year of the row (:150)
 *   workspace     dev, W2A_FEATURE_WORKSPACE_BYTES
 * Columns without accumulation are bit-identical to the reference's float64 results; rolling means agree to 1e-12. */
#define W2A_FEATURE_WORKSPACE_BYTES 4096
int w2a_engineer_features(w2a_ctx* ctx, void* stream, int64_t n_rows, int n_counties, const int64_t* county_start,
                          int max_rows_per_county, const int32_t* date_days, const double* himax_c,
                          const uint8_t* alert, const double* issued_in_advance, const uint8_t* holiday,
                          double* feat, int32_t* year, void* workspace);

/* Number of kernels this context has launched (bench.py's `gpu_launches`). */
int64_t w2a_launch_count(const w2a_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* W2A_H_ */
