/*
 * heat_oracle.c -- plain-C restatement of the reference's HeatAlertEnv step/reset for a
 * batch of envs, float64 throughout, OpenMP over envs.
 *
 * TEST INFRASTRUCTURE -- not part of the product.  Used (a) by tests as a second,
 * independently written checker next to oracle/heat_alert_oracle.py, and (b) by
 * bench.py as the `cpu_baseline` of kind "port" (the reference algorithm on the GPU
 * box's host cores).  Pinned against the unmodified reference through
 * tests/golden/reference_trajectories.npz (tests/test_oracle_golden.py).
 *
 * Follows src/weather2alert/env.py of NSAPH-Projects/weather2alert; line numbers cited
 * inline.  Tables are the unpacked ones: feat[S][Y][T][27] float64, sig[S][Y][T]
 * float32, coef[P][Sw][2][28] float32, n_days[S][Y] uint8.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define NF 27
#define OBS 29

typedef struct {
  int64_t S, Y, T, P, Sw;
  const double* feat;
  const float* sig;
  const float* coef;
  const uint8_t* n_days;
} ho_tables;

typedef struct {
  int32_t* loc;      /* [n] feature-store county (weather of the ORIGINAL location, Q7) */
  int32_t* year;     /* [n] */
  int32_t* col;      /* [n] coefficient column (Q6) */
  int32_t* p;        /* [n] posterior sample */
  int32_t* budget;   /* [n] */
  int32_t* t;        /* [n] */
  int32_t* streak;   /* [n] */
  int32_t* used;     /* [n] sum(actual_alert_buffer) */
  int32_t* n_steps;  /* [n] len(actual_alert_buffer) */
  uint8_t* last14;   /* [n][14] ring buffer: actual_alert_buffer[-14:] */
  double* obs;       /* [n][29] self.observation */
} ho_state;

static double expit(double z) { return 1.0 / (1.0 + exp(-z)); } /* scipy.special.expit, env.py:12 */

static const double* row_of(const ho_tables* tb, const ho_state* st, int64_t e, int32_t t) {
  return tb->feat + ((((int64_t)st->loc[e] * tb->Y + st->year[e]) * tb->T + t) * NF);
}

/* env.py:186-195 */
static void get_obs(const ho_tables* tb, const ho_state* st, int64_t e, double* row) {
  const int32_t t = st->t[e];
  memcpy(row, row_of(tb, st, e, t), NF * sizeof(double));
  row[27] = (double)tb->sig[((int64_t)st->loc[e] * tb->Y + st->year[e]) * tb->T + t];
  int32_t cnt = 0, last = 0;
  const int32_t n = st->n_steps[e];
  for (int k = 0; k < 14 && k < n; ++k) cnt += st->last14[e * 14 + k];
  if (n > 0) last = st->last14[e * 14 + (n - 1) % 14];
  row[23] = t > 0 ? (double)last : 0.0;            /* :190 alert_lag1 */
  row[28] = (double)cnt;                            /* :191 alert_2wks (new label, Q3) */
  row[24] = (double)st->streak[e];                  /* :192 */
  row[25] = (double)(st->budget[e] - st->used[e]);  /* :193 */
}

/* env.py:155-184 after the RNG draws; budget < -2^30 means "historical" (env.py:169) */
void ho_reset(const ho_tables* tb, ho_state* st, int64_t n, const int32_t* loc, const int32_t* year,
              const int32_t* col, const int32_t* p, const int32_t* budget, int32_t* remaining) {
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < n; ++e) {
    st->loc[e] = loc[e];
    st->year[e] = year[e];
    st->col[e] = col[e];
    st->p[e] = p[e];
    st->t[e] = 0;
    st->streak[e] = 0;
    st->used[e] = 0;
    st->n_steps[e] = 0;
    memset(st->last14 + e * 14, 0, 14);
    int32_t b = budget ? budget[e] : INT32_MIN;
    if (b < -(1 << 30)) b = (int32_t)row_of(tb, st, e, 0)[25]; /* :169 */
    st->budget[e] = b;
    get_obs(tb, st, e, st->obs + e * OBS); /* :181 */
    if (remaining) remaining[e] = b;
  }
}

/* env.py:238-262 */
void ho_step(const ho_tables* tb, ho_state* st, int64_t n, const uint8_t* action, double* reward,
             uint8_t* done, int32_t* remaining, uint8_t* at_budget) {
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < n; ++e) {
    const int a = action[e] != 0;
    const int at_b = st->used[e] == st->budget[e];  /* :242 */
    const int act = (a == 1 && at_b) ? 0 : a;       /* :243-246 */
    st->last14[e * 14 + st->n_steps[e] % 14] = (uint8_t)act; /* :248 */
    st->n_steps[e] += 1;
    st->used[e] += act;

    double row[OBS]; /* :202 */
    get_obs(tb, st, e, row);
    const float* w = tb->coef + (((int64_t)st->p[e] * tb->Sw + st->col[e]) * 2) * 28;
    double zb = 0.0, ze = 0.0;
    for (int k = 0; k < NF; ++k) { /* :206-217 */
      zb += row[k] * (double)w[k];
      ze += row[k] * (double)w[28 + k];
    }
    zb += (double)w[27]; /* bias, :203 */
    ze += (double)w[28 + 27];
    const double baseline = expit(zb);                                /* :211 */
    const double eff = expit(ze) * (row[0] > 0.5 ? 1.0 : 0.0);        /* :218 */
    reward[e] = -1000.0 / 152.0 * baseline * (1.0 - eff * (double)act); /* :221 */

    const int32_t nd = tb->n_days[(int64_t)st->loc[e] * tb->Y + st->year[e]];
    const int d = st->t[e] >= nd - 1; /* :256 */
    if (!d) {
      memcpy(st->obs + e * OBS, row, sizeof(row)); /* :258 */
      st->t[e] += 1;
      st->streak[e] = act ? st->streak[e] + 1 : 0; /* :260 */
    }
    done[e] = (uint8_t)d;
    if (remaining) remaining[e] = st->budget[e] - st->used[e];
    if (at_budget) at_budget[e] = (uint8_t)at_b;
  }
}

int ho_max_threads(void) {
#ifdef _OPENMP
  extern int omp_get_max_threads(void);
  return omp_get_max_threads();
#else
  return 1;
#endif
}
