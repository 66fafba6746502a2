// w2a_features.cuh -- the feature-store columns from raw daily weather and alert records, on the device.
//
// Replaces the pandas statements of the reference's ETL, data-processing/merge_state_actions.py:121-210
// (SURVEY.md 8(f) N4): per-county percentile rank of the heat index, hinge and interaction terms, 3/7-day
// rolling means and excess heat, day of summer, alert history (14-day count, lag, streak), weekend flag,
// summer budget / remaining budget, and the standardised cubic B-spline basis of the day of summer.
// Rows are sorted by (county, date) as the reference sorts them (:23); ranks, rolling windows, lag and streak
// run over a county's whole row sequence (they cross year boundaries, as `groupby("fips")` does), day of summer
// and budgets over (county, year) groups.
//
// One CTA per county: its rows are staged in shared memory once, a single thread makes the few inherently
// serial passes (streak, cumulative alerts, year groups: O(rows) each), every thread then finishes its rows.
// The rank is an O(rows^2) count per county (2 142 rows for 14 summers): exact average ranks without a sort.
// Arithmetic uses explicitly rounded operations in the reference's order (no FMA contraction), so that columns
// without accumulation are bit-identical to pandas' float64 results.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace w2a {

constexpr int kFeatCols = 27;       // schema.REWARD_FEATURES order (obs slots 0..26)
constexpr int kDosBins = 256;       // day-of-summer histogram (T <= 255) + 1 word for the maximum
constexpr int kFeatRowBytes = 32;   // shared memory per row of a county

struct FeatArgs {
  const long long* county_start;  // [S + 1] row offsets
  const int32_t* date_days;       // [N] days since 1970-01-01
  const double* himax_c;          // [N] daily maximum heat index, Celsius
  const uint8_t* alert;           // [N]
  const double* issued;           // [N] issued_in_advance, NaN = none
  const uint8_t* holiday;         // [N] host-supplied (the reference asks the `holidays` package, :184-187)
  double* feat;                   // [N, 27]
  int32_t* year;                  // [N]
  unsigned* hist;                 // [kDosBins + 1]: rows per day of summer, [kDosBins] = largest day of summer
  int S;
  int max_rows;                   // largest county (rows); sizes the shared memory
};

// days since 1970-01-01 -> calendar year (proleptic Gregorian; H. Hinnant's civil_from_days)
__device__ __forceinline__ int civil_year(int z) {
  z += 719468;
  const int era = (z >= 0 ? z : z - 146096) / 146097;
  const unsigned doe = static_cast<unsigned>(z - era * 146097);
  const unsigned yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
  const int y = static_cast<int>(yoe) + era * 400;
  const unsigned doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
  const unsigned mp = (5 * doy + 2) / 153;
  return y + (mp >= 10 ? 1 : 0);  // months Jan, Feb belong to the next civil year of the shifted calendar
}

__global__ void __launch_bounds__(256) feature_county_kernel(const FeatArgs p) {
  extern __shared__ __align__(16) unsigned char fsm[];
  __shared__ unsigned s_hist[kDosBins];
  __shared__ unsigned s_maxdos;
  double* const hi = reinterpret_cast<double*>(fsm);                     // hi_max
  double* const q = hi + p.max_rows;                                     // heat_qi
  int* const yr = reinterpret_cast<int*>(q + p.max_rows);                // calendar year
  int* const csum = yr + p.max_rows;                                     // alerts so far in the county (inclusive)
  int* const aux = csum + p.max_rows;                                    // streak << 16 | alert
  int* const grp = aux + p.max_rows;                                     // first row of the (county, year) group
  // (rolling alerts and the group's budget reuse csum/grp: budget = csum[last of group] - csum[grp - 1])
  for (int i = threadIdx.x; i < kDosBins; i += blockDim.x) s_hist[i] = 0u;
  if (threadIdx.x == 0) s_maxdos = 0u;
  for (int c = blockIdx.x; c < p.S; c += gridDim.x) {
    const long long r0 = p.county_start[c];
    const int n = static_cast<int>(p.county_start[c + 1] - r0);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const double h = p.himax_c[r0 + i];
      // :121  0.01 * (HImax_C * 9 / 5 + 32)
      hi[i] = __dmul_rn(0.01, __dadd_rn(__ddiv_rn(__dmul_rn(h, 9.0), 5.0), 32.0));
      yr[i] = civil_year(p.date_days[r0 + i]);
      aux[i] = p.alert[r0 + i] != 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // the serial passes: streak (:161-172), cumulative alerts, year groups
      int run = 0, cum = 0, g0 = 0;
      for (int i = 0; i < n; ++i) {
        const int a = aux[i] & 1;
        run = a ? run + 1 : 0;
        cum += a;
        if (i > 0 && yr[i] != yr[i - 1]) g0 = i;
        aux[i] = a | (run << 16);
        csum[i] = cum;
        grp[i] = g0;
      }
    }
    // :122  groupby("fips")["hi_max"].rank(pct=True): average rank of ties / number of valid values
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const double v = hi[i];
      int less = 0, equal = 0, valid = 0;
      for (int j = 0; j < n; ++j) {
        const double u = hi[j];
        less += u < v;
        equal += u == v;
        valid += u == u;
      }
      q[i] = v == v ? __ddiv_rn(static_cast<double>(less) + 0.5 * static_cast<double>(equal + 1), static_cast<double>(valid))
                    : CUDART_NAN;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const double qi = q[i], h = hi[i];
      double* f = p.feat + static_cast<size_t>(r0 + i) * kFeatCols;
      // :125-128 hinges, :131-133 interactions (note :132-133: heat_qi_above_x * hi_max)
      const double q25 = qi > 0.25 ? qi : (qi == qi ? 0.0 : qi * 0.0), q75 = qi > 0.75 ? qi : (qi == qi ? 0.0 : qi * 0.0);
      const double h25 = h > 25.0 ? h : (h == h ? 0.0 : h * 0.0), h75 = h > 75.0 ? h : (h == h ? 0.0 : h * 0.0);
      f[0] = qi; f[1] = q25; f[2] = q75; f[3] = h; f[4] = h25; f[5] = h75;
      f[6] = __dmul_rn(qi, h); f[7] = __dmul_rn(q25, h); f[8] = __dmul_rn(q75, h);
      // :136-147 rolling means over the county's rows (min_periods = 1, NaN skipped), excess heat clipped at 0
      double s3 = 0.0, s7 = 0.0;
      int c3 = 0, c7 = 0;
      for (int j = max(0, i - 6); j <= i; ++j) {
        const double u = q[j];
        if (u == u) {
          s7 = __dadd_rn(s7, u); ++c7;
          if (j >= i - 2) { s3 = __dadd_rn(s3, u); ++c3; }
        }
      }
      const double m3 = c3 ? __ddiv_rn(s3, static_cast<double>(c3)) : CUDART_NAN;
      const double m7 = c7 ? __ddiv_rn(s7, static_cast<double>(c7)) : CUDART_NAN;
      const double d3 = __dsub_rn(qi, m3), d7 = __dsub_rn(qi, m7);
      const double e3 = d3 < 0.0 ? 0.0 : d3, e7 = d7 < 0.0 ? 0.0 : d7;
      f[9] = m3; f[10] = e3; f[11] = __dmul_rn(e3, qi);
      f[12] = m7; f[13] = e7; f[14] = __dmul_rn(e7, qi);
      const int day = p.date_days[r0 + i];
      const int wd = ((day % 7) + 10) % 7;  // 1970-01-01 is a Thursday (weekday 3); Monday = 0
      f[15] = wd >= 5 ? 1.0 : 0.0;                                   // :178
      f[16] = p.holiday[r0 + i] ? 1.0 : 0.0;                         // :184-187 (host calendar)
      const int g0 = grp[i];
      const int dos = i - g0;                                        // :151 rank of the date inside (fips, year) - 1
      f[17] = static_cast<double>(dos);
      f[18] = f[19] = f[20] = 0.0;                                   // B-splines: bspline_apply_kernel
      const int a = aux[i] & 1;
      f[21] = static_cast<double>(a);                                // :114
      f[22] = static_cast<double>(csum[i] - (i >= 14 ? csum[i - 14] : 0));  // :154-156 rolling(14).sum()
      f[23] = static_cast<double>(i > 0 ? (aux[i - 1] & 1) : 0);     // :157 shift(1).fillna(0)
      f[24] = static_cast<double>(aux[i] >> 16);                     // :161-172
      int last = i;                                                  // last row of the (county, year) group
      while (last + 1 < n && grp[last + 1] == g0) ++last;
      const int before = g0 > 0 ? csum[g0 - 1] : 0;
      f[25] = static_cast<double>((csum[last] - before) - (csum[i] - before));  // :190-196 budget - rolling_alerts
      const double adv = p.issued[r0 + i];
      f[26] = adv == adv ? adv : 0.0;                                // :175 fillna(0)
      p.year[r0 + i] = yr[i];                                        // :150
      atomicAdd(&s_hist[min(dos, kDosBins - 1)], 1u);
      atomicMax(&s_maxdos, static_cast<unsigned>(dos));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kDosBins; i += blockDim.x)
    if (s_hist[i]) atomicAdd(&p.hist[i], s_hist[i]);
  if (threadIdx.x == 0) atomicMax(&p.hist[kDosBins], s_maxdos);
}

// :198-210  bs(dos / M, df=3, degree=3, lower_bound=0, upper_bound=M + 1) - 1: no interior knots, so the basis is the
// cubic Bernstein basis in u = (dos / M) / (M + 1) without its first function; each column standardised with its mean
// and (n-1) standard deviation over all rows -- both follow from the day-of-summer histogram.
__device__ __forceinline__ double bernstein3(double u, int k) {
  const double v = 1.0 - u;
  return k == 0 ? 3.0 * u * v * v : k == 1 ? 3.0 * u * u * v : u * u * u;
}

__global__ void bspline_stats_kernel(const unsigned* hist, double* stats /* mean[3], std[3], M */) {
  const int k = threadIdx.x;
  if (k >= 3) return;
  const double M = static_cast<double>(hist[kDosBins]);
  double n = 0.0, sum = 0.0;
  for (int d = 0; d < kDosBins; ++d) {
    const double c = hist[d];
    n += c;
    sum += c * bernstein3((d / M) / (M + 1.0), k);
  }
  const double mean = sum / n;
  double ss = 0.0;
  for (int d = 0; d < kDosBins; ++d) {
    const double dev = bernstein3((d / M) / (M + 1.0), k) - mean;
    ss += hist[d] * dev * dev;
  }
  stats[k] = mean;
  stats[3 + k] = sqrt(ss / (n - 1.0));
  if (k == 0) stats[6] = M;
}

__global__ void __launch_bounds__(256) bspline_apply_kernel(double* feat, const double* stats, long long n) {
  const double M = stats[6];
  for (long long r = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; r < n;
       r += static_cast<long long>(gridDim.x) * blockDim.x) {
    double* f = feat + r * kFeatCols;
    const double u = (f[17] / M) / (M + 1.0);
#pragma unroll
    for (int k = 0; k < 3; ++k) f[18 + k] = (bernstein3(u, k) - stats[k]) / stats[3 + k];
  }
}

}  // namespace w2a
