// w2a.cu -- C ABI (include/w2a.h) over the sm_100a kernels in w2a_kernels.cuh.
//
// Build (see __graft_entry__.build):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
//        -Xcompiler -fPIC -shared -I include -o weather2alert_b200/libw2a.so w2a.cu
#include "w2a.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <type_traits>
#include <vector>

#include "w2a_features.cuh"
#include "w2a_kernels.cuh"
#include "w2a_policy.cuh"

namespace {

constexpr int kHostStreams = 3;  // streams of the *_host pipeline (H2D / kernel / D2H of consecutive chunks overlap)
int64_t g_host_chunk = 0;  // envs per pipelined chunk of the *_host calls; 0 = by env-set size (W2A_HOST_CHUNK overrides: tuning)
// a quarter of the env set, between 2^16 (enough to hide the per-copy set-up) and 2^18 envs (B200: 0.93 -> 0.95 of the link at 2^20 envs)
inline int64_t host_chunk(int64_t n) {
  if (g_host_chunk > 0) return g_host_chunk;
  return std::min<int64_t>(1 << 18, std::max<int64_t>(1 << 16, (n / 4) & ~int64_t{1023}));
}

struct HostEnvs {
  int64_t n = 0;
  void* state = nullptr;
  uint8_t* action = nullptr;
  float* obs = nullptr;
  float* reward = nullptr;
  uint8_t* terminated = nullptr;
  int32_t* remaining = nullptr;
  uint8_t* at_budget = nullptr;
  int32_t* idx = nullptr;  // 5 planes of n int32: loc, year, coef_col, coef_idx, budget
  uint8_t* mask = nullptr;
  // which envs step on their own county's coefficient column (decided from the host index streams of the
  // resets): when all do and the whole-summer tables exist, the step runs the TABLE variant
  std::vector<uint8_t> noncanon;
  int64_t n_noncanon = 0;
  bool local = false;  // the last full reset's index streams were in (county, column, sample, year) order: W2A_STEP_LOCAL
  bool ever_reset = false;
  cudaStream_t streams[kHostStreams] = {};
  cudaEvent_t kernels_done[kHostStreams] = {};  // last step kernel of each stream: the small outputs leave after all of them
  // small batches (n <= kHostPackedMax): the step outputs are one device block [obs | reward | remaining |
  // terminated | at_budget] that comes back with ONE copy into a pinned staging block (and the actions go
  // out through it), whatever kind of host memory the caller's arrays are -- the per-call cost of a
  // single env is a handful of driver calls instead of six copies and three synchronisations
  bool packed = false;
  unsigned char* pack_dev = nullptr;
  unsigned char* pack_host = nullptr;  // pinned: [outputs as on the device | actions]
  size_t pack_bytes = 0;
};
constexpr int64_t kHostPackedMax = 1024;

char g_create_error[512] = "";
constexpr int kMaxKernels = 48;

}  // namespace

struct w2a_ctx {
  int device = 0;
  int sm_count = 0;
  w2a::Tables tb = {};
  bool have_features = false;
  bool have_coefs = false;
  int64_t launches = 0;
  HostEnvs* host = nullptr;
  void* table_scratch = nullptr;  // z32_all | zt_all | x32 | pt, owned by the context (w2a_precompute_logits)
  std::vector<int32_t> col_to_loc_host;  // host copy of the county map (canonical check of the *_host resets)
  int64_t keep_state_max_envs = 3 << 19;  // W2A_KEEP_STATE_MAX overrides (0 = never)
  int64_t table_rows1_max_envs = 5 << 20;  // TABLE step without the locality hint: up to this many envs one row round in flight (W2A_TABLE_ROWS1_MAX)
  int small_table_groups_per_sm = 16;  // TABLE step: env sets of up to this many 32-env groups per SM take the one-warp-per-CTA instance (W2A_SMALL_TABLE overrides, 0 = never)
  bool pdl_policy = false;        // the policy and step + policy kernels are launched serialised: as programmatic dependents they run
                                  // 1-10 us per iteration SLOWER in the graphed loop (B200, 2^20 envs; W2A_PDL_POLICY=1 to compare)
  bool pdl = true;                // step kernels are launched with programmatic stream serialization (W2A_PDL=0: off)
  const void* occ_key[kMaxKernels] = {};  // occupancy (CTAs per SM) of each kernel launched so far
  int occ_val[kMaxKernels] = {};
  int n_occ = 0;
  char err[512] = "";
};

namespace {

int fail(w2a_ctx* ctx, int code, const char* fmt, ...) {
  char* dst = ctx ? ctx->err : g_create_error;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(dst, 512, fmt, ap);
  va_end(ap);
  return code;
}

#define W2A_CUDA(ctx, expr)                                                              \
  do {                                                                                   \
    cudaError_t err__ = (expr);                                                          \
    if (err__ != cudaSuccess)                                                            \
      return fail((ctx), W2A_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(err__)); \
  } while (0)

// Launch `kernel`, optionally as a programmatic dependent launch: with the attribute set the kernel may
// become resident while the previous launch on the stream drains; the step kernels execute
// griddepcontrol.wait before they read anything, so the data dependency is kept (a no-op otherwise).
template <typename... KArgs, typename... Args>
cudaError_t launch_kernel(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, bool pdl,
                          Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(static_cast<unsigned>(block));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Persistent-style grid: enough CTAs to fill every SM at the kernel's occupancy, never
// more than the work needs; the kernels walk the envs with a grid-stride loop.
template <typename K>
int grid_for(w2a_ctx* ctx, K kernel, int block, size_t smem, int64_t work_items_per_block, int64_t n, int* grid) {
  const void* key = reinterpret_cast<const void*>(kernel);
  int per_sm = 0;
  for (int i = 0; i < ctx->n_occ; ++i)
    if (ctx->occ_key[i] == key) per_sm = ctx->occ_val[i];
  if (per_sm == 0) {  // asked once per kernel and context
    if (smem > 48 * 1024)
      W2A_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    W2A_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem));
    if (per_sm < 1) return fail(ctx, W2A_E_CUDA, "kernel does not fit on an SM (block %d, %zu B shared)", block, smem);
    if (ctx->n_occ < kMaxKernels) {
      ctx->occ_key[ctx->n_occ] = key;
      ctx->occ_val[ctx->n_occ++] = per_sm;
    }
  }
  const int64_t need = (n + work_items_per_block - 1) / work_items_per_block;
  const int64_t cap = static_cast<int64_t>(ctx->sm_count) * per_sm;
  *grid = static_cast<int>(std::max<int64_t>(1, std::min(need, cap)));
  return W2A_OK;
}

template <int ADT>
int launch_step(w2a_ctx* ctx, cudaStream_t st, const w2a::StepArgs& a, int variant, bool local) {
  int grid = 0;
  if (variant == W2A_STEP_TABLE) {
    constexpr int kBlock = 32 * w2a::kTableWarps;
    // small env sets (a step is a chain of memory latencies, not bandwidth): one warp per CTA so that the groups spread
    // over the SMs, and all eight row gathers of a group in flight together
    const int64_t groups = (static_cast<int64_t>(a.n) + 31) / 32;
    if (groups <= static_cast<int64_t>(ctx->small_table_groups_per_sm) * ctx->sm_count) {
      W2A_CUDA(ctx, launch_kernel(w2a::step_table_kernel<ADT, 8>, static_cast<int>(groups), 32,
                                  sizeof(w2a::TableScratchWarp), st, ctx->pdl, a));
    } else if (!local && a.n <= ctx->table_rows1_max_envs) {  // no locality, not yet bandwidth-bound: one row round in flight
      int rc = grid_for(ctx, w2a::step_table_kernel<ADT, 1>, kBlock, w2a::kTableSmemBytes, kBlock, a.n, &grid);
      if (rc) return rc;
      W2A_CUDA(ctx, launch_kernel(w2a::step_table_kernel<ADT, 1>, grid, kBlock, w2a::kTableSmemBytes, st, ctx->pdl, a));
    } else if (local && a.keep_state) {  // table order and the state L2-resident: no row gather before the logits are done
      int rc = grid_for(ctx, w2a::step_table_kernel<ADT, 2, true>, kBlock, w2a::kTableSmemBytes, kBlock, a.n, &grid);
      if (rc) return rc;
      W2A_CUDA(ctx, launch_kernel(w2a::step_table_kernel<ADT, 2, true>, grid, kBlock, w2a::kTableSmemBytes, st, ctx->pdl, a));
    } else {
      int rc = grid_for(ctx, w2a::step_table_kernel<ADT, 2>, kBlock, w2a::kTableSmemBytes, kBlock, a.n, &grid);
      if (rc) return rc;
      W2A_CUDA(ctx, launch_kernel(w2a::step_table_kernel<ADT, 2>, grid, kBlock, w2a::kTableSmemBytes, st, ctx->pdl, a));
    }
  } else if (variant == W2A_STEP_THREAD) {
    int rc = grid_for(ctx, w2a::step_thread_kernel<ADT>, 128, 0, 128, a.n, &grid);
    if (rc) return rc;
    W2A_CUDA(ctx, launch_kernel(w2a::step_thread_kernel<ADT>, grid, 128, 0, st, ctx->pdl, a));
  } else {
    constexpr int kBlock = 32 * w2a::kCoopWarps;
    int rc = grid_for(ctx, w2a::step_coop_kernel<ADT>, kBlock, w2a::kCoopSmemBytes, kBlock, a.n, &grid);
    if (rc) return rc;
    // small batches (launch-latency bound): spread the warps over the SMs instead of packing 8 per CTA
    const int64_t groups = (static_cast<int64_t>(a.n) + 31) / 32;
    const int warps = static_cast<int>(std::min<int64_t>(w2a::kCoopWarps, std::max<int64_t>(1, groups / ctx->sm_count)));
    if (warps <= 2) {  // at most two groups per SM: every round of a group in flight at once, one or two warps per CTA
      int unused = 0;  // (sets the instance's shared-memory attribute once)
      rc = grid_for(ctx, w2a::step_coop_kernel<ADT, 7>, kBlock, w2a::kCoopSmemBytes, kBlock, a.n, &unused);
      if (rc) return rc;
      const int small_grid = static_cast<int>((groups + warps - 1) / warps);
      W2A_CUDA(ctx, launch_kernel(w2a::step_coop_kernel<ADT, 7>, small_grid, 32 * warps, warps * sizeof(w2a::WarpScratch), st,
                                  ctx->pdl, a));
    } else {
      W2A_CUDA(ctx, launch_kernel(w2a::step_coop_kernel<ADT>, grid, kBlock, w2a::kCoopSmemBytes, st, ctx->pdl, a));
    }
  }
  ctx->launches++;
  W2A_CUDA(ctx, cudaGetLastError());
  return W2A_OK;
}

int step_planes(w2a_ctx* ctx, cudaStream_t st, int64_t n, const void* action, int action_dtype,
                const uint2* st_static, uint2* st_mut, float* obs, float* reward,
                uint8_t* terminated, int32_t* remaining, uint8_t* at_budget, int variant) {
  const bool local = (variant & W2A_STEP_LOCAL) != 0;  // tuning hint, see w2a.h
  variant &= ~W2A_STEP_LOCAL;
  if (variant == W2A_STEP_AUTO) variant = W2A_STEP_COOP;  // faster than THREAD at every batch size, n = 1 included
  if (variant != W2A_STEP_THREAD && variant != W2A_STEP_COOP && variant != W2A_STEP_TABLE)
    return fail(ctx, W2A_E_INVALID, "w2a_step: unknown variant %d", variant);
  if (variant == W2A_STEP_TABLE && !ctx->tb.z32_all)
    return fail(ctx, W2A_E_NOTREADY, "w2a_step: W2A_STEP_TABLE needs w2a_set_county_map + w2a_precompute_logits first");
  w2a::StepArgs a;
  a.tb = ctx->tb;
  a.st_static = st_static;
  a.st_mut = st_mut;
  a.action = action;
  a.obs = obs;
  a.reward = reward;
  a.terminated = terminated;
  a.remaining = remaining;
  a.at_budget = at_budget;
  a.n = static_cast<int>(n);
  // 16 B of state per env are read and 8 B rewritten every step: up to 1.5M envs (24 MB of the 126 MB L2) they are
  // kept L2-resident between steps (evict_last), beyond that they stream
  a.keep_state = n <= ctx->keep_state_max_envs ? 1 : 0;
  switch (action_dtype) {
    case W2A_ACT_U8: return launch_step<0>(ctx, st, a, variant, local);
    case W2A_ACT_I32: return launch_step<1>(ctx, st, a, variant, local);
    case W2A_ACT_I64: return launch_step<2>(ctx, st, a, variant, local);
    default: return fail(ctx, W2A_E_INVALID, "w2a_step: unknown action_dtype %d", action_dtype);
  }
}

int reset_planes(w2a_ctx* ctx, cudaStream_t st, int64_t n, const int32_t* loc, const int32_t* year,
                 const int32_t* coef_col, const int32_t* coef_idx, const int32_t* budget,
                 const uint8_t* mask, uint2* st_static, uint2* st_mut, float* obs,
                 int32_t* remaining) {
  w2a::ResetArgs a;
  a.tb = ctx->tb;
  a.loc = loc;
  a.year = year;
  a.coef_col = coef_col;
  a.coef_idx = coef_idx;
  a.budget = budget;
  a.mask = mask;
  a.st_static = st_static;
  a.st_mut = st_mut;
  a.obs = obs;
  a.remaining = remaining;
  a.n = static_cast<int>(n);
  int grid = 0;
  if (ctx->tb.x32) {  // the float32 observation rows exist: one 128-byte gather per env, bulk stores
    constexpr int kBlock = 32 * w2a::kTableWarps;
    int rc = grid_for(ctx, w2a::reset_table_kernel, kBlock, w2a::kTableSmemBytes, kBlock, n, &grid);
    if (rc) return rc;
    w2a::reset_table_kernel<<<grid, kBlock, w2a::kTableSmemBytes, st>>>(a);
  } else {
    int rc = grid_for(ctx, w2a::reset_kernel, 256, 0, 256, n, &grid);
    if (rc) return rc;
    w2a::reset_kernel<<<grid, 256, 0, st>>>(a);
  }
  ctx->launches++;
  W2A_CUDA(ctx, cudaGetLastError());
  return W2A_OK;
}

int check_ready(w2a_ctx* ctx, const char* who, int64_t n) {
  if (!ctx) return W2A_E_INVALID;
  if (!ctx->have_features || !ctx->have_coefs)
    return fail(ctx, W2A_E_NOTREADY, "%s: call w2a_set_features and w2a_set_coefs first", who);
  if (n < 0 || n >= (int64_t{1} << 31) - 64)
    return fail(ctx, W2A_E_INVALID, "%s: n=%lld out of range", who, static_cast<long long>(n));
  return W2A_OK;
}

inline uint2* plane_static(void* state) { return reinterpret_cast<uint2*>(state); }
inline uint2* plane_mut(void* state, int64_t n) { return reinterpret_cast<uint2*>(state) + n; }

}  // namespace

extern "C" {

int w2a_version(void) { return W2A_VERSION; }

const char* w2a_last_error(const w2a_ctx* ctx) { return ctx ? ctx->err : g_create_error; }

int w2a_create(int device, w2a_ctx** out) {
  if (!out) return fail(nullptr, W2A_E_INVALID, "w2a_create: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess)
    return fail(nullptr, W2A_E_CUDA, "w2a_create: cudaGetDeviceCount: %s", cudaGetErrorString(e));
  if (device < 0 || device >= count)
    return fail(nullptr, W2A_E_INVALID, "w2a_create: device %d not in [0,%d)", device, count);
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess)
    return fail(nullptr, W2A_E_CUDA, "w2a_create: cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, W2A_E_INVALID,
                "w2a_create: device %d is sm_%d%d; this library carries sm_100a code only", device,
                prop.major, prop.minor);
  w2a_ctx* ctx = new (std::nothrow) w2a_ctx();
  if (!ctx) return fail(nullptr, W2A_E_NOMEM, "w2a_create: out of host memory");
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  const char* pdl_env = std::getenv("W2A_PDL");  // tuning switch: W2A_PDL=0 launches the step kernels fully serialised
  ctx->pdl = !(pdl_env && pdl_env[0] == '0');
  if (const char* ks = std::getenv("W2A_KEEP_STATE_MAX")) ctx->keep_state_max_envs = std::atoll(ks);
  if (const char* r1 = std::getenv("W2A_TABLE_ROWS1_MAX")) ctx->table_rows1_max_envs = std::atoll(r1);
  if (const char* sm = std::getenv("W2A_SMALL_TABLE")) ctx->small_table_groups_per_sm = std::atoi(sm);
  if (const char* pp = std::getenv("W2A_PDL_POLICY")) ctx->pdl_policy = std::atoi(pp) != 0;
  if (const char* hc = std::getenv("W2A_HOST_CHUNK")) g_host_chunk = std::max<int64_t>(1024, std::atoll(hc));
  *out = ctx;
  return W2A_OK;
}

int w2a_destroy(w2a_ctx* ctx) {
  if (!ctx) return W2A_OK;
  w2a_host_envs_destroy(ctx);
  if (ctx->table_scratch) {
    cudaSetDevice(ctx->device);
    cudaFree(ctx->table_scratch);
    ctx->table_scratch = nullptr;
  }
  delete ctx;
  return W2A_OK;
}

// the whole-summer tables belong to one (features, coefficients, county map) triple
static void drop_tables(w2a_ctx* ctx) {
  if (ctx->table_scratch) {
    cudaSetDevice(ctx->device);
    cudaFree(ctx->table_scratch);
    ctx->table_scratch = nullptr;
  }
  ctx->tb.z32_all = nullptr;
  ctx->tb.zt_all = nullptr;
  ctx->tb.x32 = nullptr;
  ctx->tb.pt = nullptr;
  ctx->tb.n_records = 0;
}

int w2a_set_features(w2a_ctx* ctx, const double* x, const uint8_t* n_days, const int32_t* budget0, int S, int Y,
                     int T) {
  if (!ctx) return W2A_E_INVALID;
  if (!x || !n_days || !budget0) return fail(ctx, W2A_E_INVALID, "w2a_set_features: NULL table");
  if (S <= 0 || Y <= 0 || T <= 0 || T > 255 || Y > 255)
    return fail(ctx, W2A_E_INVALID, "w2a_set_features: bad shape S=%d Y=%d T=%d (Y<=255, T<=255)", S, Y, T);
  if (static_cast<int64_t>(S) * Y * T >= (int64_t{1} << 24))
    return fail(ctx, W2A_E_INVALID, "w2a_set_features: S*Y*T=%lld rows exceed the 24-bit row index",
                static_cast<long long>(static_cast<int64_t>(S) * Y * T));
  if (reinterpret_cast<uintptr_t>(x) & 127)
    return fail(ctx, W2A_E_INVALID, "w2a_set_features: the table must be 128-byte aligned");
  ctx->tb.x = x;
  ctx->tb.n_days = n_days;
  ctx->tb.budget0 = budget0;
  ctx->tb.S = S;
  ctx->tb.Y = Y;
  ctx->tb.T = T;
  ctx->have_features = true;
  ctx->tb.col_to_loc = nullptr;  // a county map (and the tables built from it) belongs to one set of tables
  ctx->col_to_loc_host.clear();
  drop_tables(ctx);
  return W2A_OK;
}

int w2a_set_coefs(w2a_ctx* ctx, const float* w, int P, int S_w) {
  if (!ctx) return W2A_E_INVALID;
  if (!w || P <= 0 || S_w <= 0) return fail(ctx, W2A_E_INVALID, "w2a_set_coefs: bad argument");
  if (reinterpret_cast<uintptr_t>(w) & 127)
    return fail(ctx, W2A_E_INVALID, "w2a_set_coefs: the table must be 128-byte aligned");
  if (static_cast<int64_t>(P) * S_w >= (int64_t{1} << 24))
    return fail(ctx, W2A_E_INVALID, "w2a_set_coefs: P*S_w=%lld records exceed the 24-bit record index",
                static_cast<long long>(static_cast<int64_t>(P) * S_w));
  ctx->tb.w = w;
  ctx->tb.P = P;
  ctx->tb.Sw = S_w;
  ctx->have_coefs = true;
  ctx->tb.col_to_loc = nullptr;
  ctx->col_to_loc_host.clear();
  drop_tables(ctx);
  return W2A_OK;
}

int w2a_set_county_map(w2a_ctx* ctx, const int32_t* col_to_loc) {
  if (!ctx) return W2A_E_INVALID;
  if (!ctx->have_features || !ctx->have_coefs)
    return fail(ctx, W2A_E_NOTREADY, "w2a_set_county_map: call w2a_set_features and w2a_set_coefs first");
  if (!col_to_loc) return fail(ctx, W2A_E_INVALID, "w2a_set_county_map: NULL map");
  W2A_CUDA(ctx, cudaSetDevice(ctx->device));
  drop_tables(ctx);
  ctx->col_to_loc_host.resize(static_cast<size_t>(ctx->tb.Sw));
  W2A_CUDA(ctx, cudaMemcpy(ctx->col_to_loc_host.data(), col_to_loc, sizeof(int32_t) * ctx->tb.Sw, cudaMemcpyDeviceToHost));
  for (int32_t s : ctx->col_to_loc_host)
    if (s < -1 || s >= ctx->tb.S) {
      ctx->col_to_loc_host.clear();
      return fail(ctx, W2A_E_INVALID, "w2a_set_county_map: county %d not in [-1, %d)", s, ctx->tb.S);
    }
  ctx->tb.col_to_loc = col_to_loc;
  return W2A_OK;
}

namespace {
struct TableBytes {
  size_t z32, zt, x32, pt;
  size_t total() const { return z32 + zt + x32 + pt; }
};
TableBytes table_bytes(const w2a::Tables& tb) {
  const size_t entries = static_cast<size_t>(tb.T) * tb.Sw * tb.P * tb.Y;
  const size_t rows = static_cast<size_t>(tb.S) * tb.Y * tb.T;
  return {entries * 32, entries * 16, rows * 32 * sizeof(float), (rows * sizeof(double) + 127) & ~size_t{127}};
}
}  // namespace

size_t w2a_precompute_bytes(const w2a_ctx* ctx) {
  if (!ctx || !ctx->have_features || !ctx->have_coefs) return 0;
  return table_bytes(ctx->tb).total();
}

int w2a_precompute_logits(w2a_ctx* ctx, void* stream) {
  if (!ctx) return W2A_E_INVALID;
  if (!ctx->have_features || !ctx->have_coefs || !ctx->tb.col_to_loc)
    return fail(ctx, W2A_E_NOTREADY, "w2a_precompute_logits: call w2a_set_features, w2a_set_coefs and w2a_set_county_map first");
  W2A_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  drop_tables(ctx);
  w2a::Tables& tb = ctx->tb;
  const int64_t entries = static_cast<int64_t>(tb.Sw) * tb.P * tb.Y * tb.T;
  if (entries >= (int64_t{1} << 32))
    return fail(ctx, W2A_E_INVALID, "w2a_precompute_logits: %lld table entries exceed the 32-bit entry index",
                static_cast<long long>(entries));
  const TableBytes b = table_bytes(tb);
  size_t free_b = 0, total_b = 0;
  W2A_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
  if (b.total() + (size_t{1} << 30) > free_b)
    return fail(ctx, W2A_E_NOMEM, "w2a_precompute_logits: the tables need %.2f GB, %.2f GB are free", b.total() * 1e-9,
                free_b * 1e-9);
  W2A_CUDA(ctx, cudaMalloc(&ctx->table_scratch, b.total()));
  char* base = static_cast<char*>(ctx->table_scratch);
  float4* z32_all = reinterpret_cast<float4*>(base);
  double2* zt_all = reinterpret_cast<double2*>(base + b.z32);
  float* x32 = reinterpret_cast<float*>(base + b.z32 + b.zt);
  double* pt = reinterpret_cast<double*>(base + b.z32 + b.zt + b.x32);
  W2A_CUDA(ctx, cudaMemsetAsync(base, 0, b.z32 + b.zt, st));  // columns without a county keep zero entries
  tb.n_records = static_cast<unsigned>(tb.Sw) * static_cast<unsigned>(tb.P);
  const size_t rows = static_cast<size_t>(tb.S) * tb.Y * tb.T;
  {
    const int grid = static_cast<int>(std::min<size_t>((rows + 7) / 8, static_cast<size_t>(ctx->sm_count) * 16));
    w2a::obs_rows_kernel<<<grid, 256, 0, st>>>(tb.x, x32, static_cast<unsigned>(rows));
    ctx->launches++;
  }
  {
    const int tasks = tb.Sw * ((tb.P + 7) / 8);
    const unsigned gy = static_cast<unsigned>((tasks + w2a::kTableTaskWarps - 1) / w2a::kTableTaskWarps);
    if (gy > 65535u) {
      cudaFree(ctx->table_scratch);
      ctx->table_scratch = nullptr;
      return fail(ctx, W2A_E_INVALID, "w2a_precompute_logits: %d (column, sample block) tasks exceed the launch grid", tasks);
    }
    // blockIdx.x = day, blockIdx.y = block of tasks
    w2a::logit_all_kernel<<<dim3(static_cast<unsigned>(tb.T), gy), 32 * w2a::kTableTaskWarps, 0, st>>>(tb, zt_all, z32_all, 0);
    ctx->launches++;
  }
  W2A_CUDA(ctx, cudaGetLastError());
  tb.z32_all = z32_all;
  tb.zt_all = zt_all;
  tb.x32 = x32;
  tb.pt = pt;
  return W2A_OK;
}

size_t w2a_state_bytes(int64_t n) { return n < 0 ? 0 : static_cast<size_t>(n) * 16; }

int w2a_reset(w2a_ctx* ctx, void* stream, int64_t n, const int32_t* loc, const int32_t* year,
              const int32_t* coef_col, const int32_t* coef_idx, const int32_t* budget,
              const uint8_t* mask, void* state, float* obs, int32_t* remaining) {
  int rc = check_ready(ctx, "w2a_reset", n);
  if (rc) return rc;
  if (n == 0) return W2A_OK;
  if (!loc || !year || !coef_col || !coef_idx || !state || !obs)
    return fail(ctx, W2A_E_INVALID, "w2a_reset: NULL argument");
  W2A_CUDA(ctx, cudaSetDevice(ctx->device));
  return reset_planes(ctx, static_cast<cudaStream_t>(stream), n, loc, year, coef_col, coef_idx, budget,
                      mask, plane_static(state), plane_mut(state, n), obs, remaining);
}

int w2a_step(w2a_ctx* ctx, void* stream, int64_t n, const void* action, int action_dtype, void* state,
             float* obs, float* reward, uint8_t* terminated, int32_t* remaining, uint8_t* at_budget,
             int variant) {
  int rc = check_ready(ctx, "w2a_step", n);
  if (rc) return rc;
  if (n == 0) return W2A_OK;
  if (!action || !state || !obs || !reward || !terminated)
    return fail(ctx, W2A_E_INVALID, "w2a_step: NULL argument");
  W2A_CUDA(ctx, cudaSetDevice(ctx->device));
  return step_planes(ctx, static_cast<cudaStream_t>(stream), n, action, action_dtype, plane_static(state),
                     plane_mut(state, n), obs, reward, terminated, remaining, at_budget, variant);
}

// ---- host-buffer entry points ---------------------------------------------------------------

int w2a_host_envs_destroy(w2a_ctx* ctx) {
  if (!ctx || !ctx->host) return W2A_OK;
  HostEnvs* h = ctx->host;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  cudaFree(h->state);
  cudaFree(h->action);
  if (h->packed) {
    cudaFree(h->pack_dev);
    cudaFreeHost(h->pack_host);
  } else {
    cudaFree(h->obs);
    cudaFree(h->reward);
    cudaFree(h->terminated);
    cudaFree(h->remaining);
    cudaFree(h->at_budget);
  }
  cudaFree(h->idx);
  cudaFree(h->mask);
  for (auto& s : h->streams)
    if (s) cudaStreamDestroy(s);
  for (auto& e : h->kernels_done)
    if (e) cudaEventDestroy(e);
  delete h;
  ctx->host = nullptr;
  return W2A_OK;
}

int w2a_host_envs_create(w2a_ctx* ctx, int64_t n) {
  int rc = check_ready(ctx, "w2a_host_envs_create", n);
  if (rc) return rc;
  if (n <= 0) return fail(ctx, W2A_E_INVALID, "w2a_host_envs_create: n must be positive");
  w2a_host_envs_destroy(ctx);
  W2A_CUDA(ctx, cudaSetDevice(ctx->device));
  HostEnvs* h = new (std::nothrow) HostEnvs();
  if (!h) return fail(ctx, W2A_E_NOMEM, "w2a_host_envs_create: out of host memory");
  ctx->host = h;
  h->n = n;
  const size_t un = static_cast<size_t>(n);
  W2A_CUDA(ctx, cudaMalloc(&h->state, w2a_state_bytes(n)));
  W2A_CUDA(ctx, cudaMalloc(&h->action, un));
  h->packed = n <= kHostPackedMax;
  if (h->packed) {
    const size_t obs_b = un * W2A_OBS_DIM * sizeof(float);
    h->pack_bytes = obs_b + un * 4 + un * 4 + un + un;
    W2A_CUDA(ctx, cudaMalloc(&h->pack_dev, h->pack_bytes));
    W2A_CUDA(ctx, cudaMallocHost(&h->pack_host, h->pack_bytes + un));
    h->obs = reinterpret_cast<float*>(h->pack_dev);
    h->reward = reinterpret_cast<float*>(h->pack_dev + obs_b);
    h->remaining = reinterpret_cast<int32_t*>(h->pack_dev + obs_b + un * 4);
    h->terminated = h->pack_dev + obs_b + un * 8;
    h->at_budget = h->pack_dev + obs_b + un * 9;
  } else {
    W2A_CUDA(ctx, cudaMalloc(&h->obs, un * W2A_OBS_DIM * sizeof(float)));
    W2A_CUDA(ctx, cudaMalloc(&h->reward, un * sizeof(float)));
    W2A_CUDA(ctx, cudaMalloc(&h->terminated, un));
    W2A_CUDA(ctx, cudaMalloc(&h->remaining, un * sizeof(int32_t)));
    W2A_CUDA(ctx, cudaMalloc(&h->at_budget, un));
  }
  W2A_CUDA(ctx, cudaMalloc(&h->idx, un * 5 * sizeof(int32_t)));
  W2A_CUDA(ctx, cudaMalloc(&h->mask, un));
  W2A_CUDA(ctx, cudaMemset(h->obs, 0, un * W2A_OBS_DIM * sizeof(float)));
  h->noncanon.assign(un, 1);
  h->n_noncanon = n;
  for (auto& s : h->streams) W2A_CUDA(ctx, cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  for (auto& e : h->kernels_done) W2A_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return W2A_OK;
}

int w2a_reset_host(w2a_ctx* ctx, const int32_t* loc, const int32_t* year, const int32_t* coef_col,
                   const int32_t* coef_idx, const int32_t* budget, const uint8_t* mask, float* obs,
                   int32_t* remaining) {
  if (!ctx) return W2A_E_INVALID;
  HostEnvs* h = ctx->host;
  if (!h) return fail(ctx, W2A_E_NOTREADY, "w2a_reset_host: call w2a_host_envs_create first");
  if (!loc || !year || !coef_col || !coef_idx || !obs)
    return fail(ctx, W2A_E_INVALID, "w2a_reset_host: NULL argument");
  W2A_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t n = h->n;
  uint2* sstat = plane_static(h->state);
  uint2* smut = plane_mut(h->state, n);
  // an env is canonical when its coefficient column is its own county: known here from the host streams
  {
    const std::vector<int32_t>& map = ctx->col_to_loc_host;
    for (int64_t i = 0; i < n; ++i) {
      if (mask && !mask[i]) continue;
      const int32_t c = coef_col[i];
      const uint8_t non = !(c >= 0 && static_cast<size_t>(c) < map.size() && map[static_cast<size_t>(c)] == loc[i]);
      h->n_noncanon += static_cast<int64_t>(non) - static_cast<int64_t>(h->noncanon[static_cast<size_t>(i)]);
      h->noncanon[static_cast<size_t>(i)] = non;
    }
    // env slots in table order? (a full reset decides; a masked one scatters fresh episodes over them)
    if (mask) {
      h->local = false;
    } else {
      bool sorted = true;
      for (int64_t i = 1; i < n && sorted; ++i) {
        if (loc[i] != loc[i - 1]) sorted = loc[i] > loc[i - 1];
        else if (coef_col[i] != coef_col[i - 1]) sorted = coef_col[i] > coef_col[i - 1];
        else if (coef_idx[i] != coef_idx[i - 1]) sorted = coef_idx[i] > coef_idx[i - 1];
        else sorted = year[i] >= year[i - 1];
      }
      h->local = sorted;
    }
  }
  int ci = 0;
  const int64_t chunk = host_chunk(n);
  for (int64_t c0 = 0; c0 < n; c0 += chunk, ++ci) {
    const int64_t m = std::min(chunk, n - c0);
    const size_t um = static_cast<size_t>(m);
    cudaStream_t st = h->streams[ci % kHostStreams];
    int32_t* d_loc = h->idx + c0;
    int32_t* d_year = h->idx + n + c0;
    int32_t* d_col = h->idx + 2 * n + c0;
    int32_t* d_ci = h->idx + 3 * n + c0;
    int32_t* d_bud = budget ? h->idx + 4 * n + c0 : nullptr;
    uint8_t* d_mask = mask ? h->mask + c0 : nullptr;
    W2A_CUDA(ctx, cudaMemcpyAsync(d_loc, loc + c0, um * 4, cudaMemcpyHostToDevice, st));
    W2A_CUDA(ctx, cudaMemcpyAsync(d_year, year + c0, um * 4, cudaMemcpyHostToDevice, st));
    W2A_CUDA(ctx, cudaMemcpyAsync(d_col, coef_col + c0, um * 4, cudaMemcpyHostToDevice, st));
    W2A_CUDA(ctx, cudaMemcpyAsync(d_ci, coef_idx + c0, um * 4, cudaMemcpyHostToDevice, st));
    if (budget) W2A_CUDA(ctx, cudaMemcpyAsync(d_bud, budget + c0, um * 4, cudaMemcpyHostToDevice, st));
    if (mask) W2A_CUDA(ctx, cudaMemcpyAsync(d_mask, mask + c0, um, cudaMemcpyHostToDevice, st));
    int rc = reset_planes(ctx, st, m, d_loc, d_year, d_col, d_ci, d_bud, d_mask, sstat + c0, smut + c0,
                          h->obs + c0 * W2A_OBS_DIM, h->remaining + c0);
    if (rc) return rc;
    W2A_CUDA(ctx, cudaMemcpyAsync(obs + c0 * W2A_OBS_DIM, h->obs + c0 * W2A_OBS_DIM,
                                  um * W2A_OBS_DIM * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (remaining)
      W2A_CUDA(ctx, cudaMemcpyAsync(remaining + c0, h->remaining + c0, um * 4, cudaMemcpyDeviceToHost, st));
  }
  for (auto& s : h->streams) W2A_CUDA(ctx, cudaStreamSynchronize(s));
  return W2A_OK;
}

int w2a_step_host(w2a_ctx* ctx, const uint8_t* action, float* obs, float* reward, uint8_t* terminated,
                  int32_t* remaining, uint8_t* at_budget) {
  if (!ctx) return W2A_E_INVALID;
  HostEnvs* h = ctx->host;
  if (!h) return fail(ctx, W2A_E_NOTREADY, "w2a_step_host: call w2a_host_envs_create first");
  if (!action || !obs || !reward || !terminated)
    return fail(ctx, W2A_E_INVALID, "w2a_step_host: NULL argument");
  W2A_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t n = h->n;
  uint2* sstat = plane_static(h->state);
  uint2* smut = plane_mut(h->state, n);
  // one kernel variant for the whole batch, so that chunking never changes a result bit: the tabulated
  // logits when every env steps on its own county's column and the tables exist, else the cooperative kernel
  const int variant = (ctx->tb.z32_all && h->n_noncanon == 0) ? (W2A_STEP_TABLE | (h->local ? W2A_STEP_LOCAL : 0)) : W2A_STEP_COOP;
  if (h->packed) {
    const size_t un = static_cast<size_t>(n), obs_b = un * W2A_OBS_DIM * sizeof(float);
    cudaStream_t st = h->streams[0];
    unsigned char* stage_act = h->pack_host + h->pack_bytes;
    std::memcpy(stage_act, action, un);
    W2A_CUDA(ctx, cudaMemcpyAsync(h->action, stage_act, un, cudaMemcpyHostToDevice, st));
    int rc = step_planes(ctx, st, n, h->action, W2A_ACT_U8, sstat, smut, h->obs, h->reward, h->terminated, h->remaining,
                         h->at_budget, variant);
    if (rc) return rc;
    W2A_CUDA(ctx, cudaMemcpyAsync(h->pack_host, h->pack_dev, h->pack_bytes, cudaMemcpyDeviceToHost, st));
    W2A_CUDA(ctx, cudaStreamSynchronize(st));
    std::memcpy(obs, h->pack_host, obs_b);
    std::memcpy(reward, h->pack_host + obs_b, un * 4);
    if (remaining) std::memcpy(remaining, h->pack_host + obs_b + un * 4, un * 4);
    std::memcpy(terminated, h->pack_host + obs_b + un * 8, un);
    if (at_budget) std::memcpy(at_budget, h->pack_host + obs_b + un * 9, un);
    return W2A_OK;
  }
  int ci = 0;
  const int64_t chunk = host_chunk(n);
  for (int64_t c0 = 0; c0 < n; c0 += chunk, ++ci) {
    const int64_t m = std::min(chunk, n - c0);
    const size_t um = static_cast<size_t>(m);
    cudaStream_t st = h->streams[ci % kHostStreams];
    W2A_CUDA(ctx, cudaMemcpyAsync(h->action + c0, action + c0, um, cudaMemcpyHostToDevice, st));
    int rc = step_planes(ctx, st, m, h->action + c0, W2A_ACT_U8, sstat + c0, smut + c0,
                         h->obs + c0 * W2A_OBS_DIM, h->reward + c0, h->terminated + c0,
                         remaining ? h->remaining + c0 : nullptr, at_budget ? h->at_budget + c0 : nullptr,
                         variant);
    if (rc) return rc;
    W2A_CUDA(ctx, cudaEventRecord(h->kernels_done[ci % kHostStreams], st));
    // the observations (92 % of the bytes) leave chunk by chunk, behind their kernel ...
    W2A_CUDA(ctx, cudaMemcpyAsync(obs + c0 * W2A_OBS_DIM, h->obs + c0 * W2A_OBS_DIM,
                                  um * W2A_OBS_DIM * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  // ... the four small outputs as ONE copy each, once every chunk's kernel has run: 4 DMA set-ups per step instead of
  // 4 per chunk (64 at 2^20 envs), which the link spent ~5 % of a step on
  cudaStream_t tail = h->streams[0];
  for (int i = 1; i < std::min(ci, kHostStreams); ++i) W2A_CUDA(ctx, cudaStreamWaitEvent(tail, h->kernels_done[i], 0));
  const size_t un = static_cast<size_t>(n);
  W2A_CUDA(ctx, cudaMemcpyAsync(reward, h->reward, un * 4, cudaMemcpyDeviceToHost, tail));
  W2A_CUDA(ctx, cudaMemcpyAsync(terminated, h->terminated, un, cudaMemcpyDeviceToHost, tail));
  if (remaining) W2A_CUDA(ctx, cudaMemcpyAsync(remaining, h->remaining, un * 4, cudaMemcpyDeviceToHost, tail));
  if (at_budget) W2A_CUDA(ctx, cudaMemcpyAsync(at_budget, h->at_budget, un, cudaMemcpyDeviceToHost, tail));
  for (auto& s : h->streams) W2A_CUDA(ctx, cudaStreamSynchronize(s));
  return W2A_OK;
}

namespace {
int launch_rollout(w2a_ctx* ctx, void* stream, int64_t n, int steps, w2a::RolloutArgs& a, void* state, double* ep_return,
                   double* stats, bool policy) {
  W2A_CUDA(ctx, cudaSetDevice(ctx->device));
  a.tb = ctx->tb;
  a.st_static = plane_static(state);
  a.st_mut = plane_mut(state, n);
  a.ep_return = ep_return;
  a.stats = stats;
  a.n = static_cast<int>(n);
  a.steps = steps;
  int grid = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // with the whole-summer tables the kernels stream 16 B of tabulated logits per env-step (envs on a similar
  // county's column fall back to the day-rows inside the same launch)
  const bool tab = ctx->tb.zt_all != nullptr;
  auto go = [&](auto kernel) {
    int rc = grid_for(ctx, kernel, 128, 0, 16, n, &grid);  // 4 warps x 4 envs per CTA pass
    if (rc) return rc;
    kernel<<<grid, 128, 0, st>>>(a);
    return W2A_OK;
  };
  if (policy && tab) {
    const unsigned rows = static_cast<unsigned>(ctx->tb.S) * ctx->tb.Y * ctx->tb.T;
    w2a::policy_rows_kernel<<<std::min<unsigned>((rows + 255) / 256, ctx->sm_count * 8), 256, 0, st>>>(ctx->tb.x, a.policy,
                                                                                                      ctx->tb.pt, rows);
    ctx->launches++;
  }
  int rc = W2A_OK;
  if (tab) {  // table launch for the canonical envs, direct launch for the rest (each env belongs to exactly one)
    a.only = 1;
    rc = policy ? go(w2a::rollout_kernel<true, true>) : go(w2a::rollout_kernel<false, true>);
    if (rc) return rc;
    ctx->launches++;
    a.only = 2;
  }
  rc = policy ? go(w2a::rollout_kernel<true, false>) : go(w2a::rollout_kernel<false, false>);
  if (rc) return rc;
  ctx->launches++;
  W2A_CUDA(ctx, cudaGetLastError());
  return W2A_OK;
}
}  // namespace

int w2a_rollout(w2a_ctx* ctx, void* stream, int64_t n, int steps, const uint8_t* actions, uint64_t seed, float p_alert,
                void* state, double* ep_return, double* stats) {
  int rc = check_ready(ctx, "w2a_rollout", n);
  if (rc) return rc;
  if (n == 0 || steps == 0) return W2A_OK;
  if (!state || !ep_return || steps < 0) return fail(ctx, W2A_E_INVALID, "w2a_rollout: bad argument");
  if (!actions && !(p_alert >= 0.f && p_alert <= 1.f))
    return fail(ctx, W2A_E_INVALID, "w2a_rollout: p_alert=%g not in [0,1]", static_cast<double>(p_alert));
  w2a::RolloutArgs a{};
  a.actions = actions;
  a.seed = seed;
  const double thr = static_cast<double>(p_alert) * 4294967296.0;
  a.threshold = thr >= 4294967295.0 ? 0xffffffffu : static_cast<unsigned>(thr);
  return launch_rollout(ctx, stream, n, steps, a, state, ep_return, stats, false);
}

int w2a_rollout_policy(w2a_ctx* ctx, void* stream, int64_t n, int steps, const double* policy, int stochastic,
                       uint64_t seed, void* state, double* ep_return, double* stats) {
  int rc = check_ready(ctx, "w2a_rollout_policy", n);
  if (rc) return rc;
  if (n == 0 || steps == 0) return W2A_OK;
  if (!state || !ep_return || !policy || steps < 0)
    return fail(ctx, W2A_E_INVALID, "w2a_rollout_policy: bad argument");
  w2a::RolloutArgs a{};
  a.policy = policy;
  a.stochastic = stochastic != 0;
  a.seed = seed;
  return launch_rollout(ctx, stream, n, steps, a, state, ep_return, stats, true);
}

int w2a_engineer_features(w2a_ctx* ctx, void* stream, int64_t n_rows, int n_counties, const int64_t* county_start,
                          int max_rows_per_county, const int32_t* date_days, const double* himax_c, const uint8_t* alert,
                          const double* issued_in_advance, const uint8_t* holiday, double* feat, int32_t* year,
                          void* workspace) {
  if (!ctx) return W2A_E_INVALID;
  if (n_rows == 0 || n_counties == 0) return W2A_OK;
  if (n_rows < 0 || n_counties < 0 || !county_start || !date_days || !himax_c || !alert || !issued_in_advance || !holiday ||
      !feat || !year || !workspace)
    return fail(ctx, W2A_E_INVALID, "w2a_engineer_features: NULL or negative argument");
  const size_t smem = static_cast<size_t>(max_rows_per_county) * w2a::kFeatRowBytes;
  if (max_rows_per_county <= 0 || smem > 200 * 1024)
    return fail(ctx, W2A_E_INVALID, "w2a_engineer_features: max_rows_per_county=%d not in [1, %d] (one county is staged in shared memory)",
                max_rows_per_county, 200 * 1024 / w2a::kFeatRowBytes);
  W2A_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned* hist = static_cast<unsigned*>(workspace);
  double* stats = reinterpret_cast<double*>(static_cast<char*>(workspace) + 2048);
  W2A_CUDA(ctx, cudaMemsetAsync(workspace, 0, W2A_FEATURE_WORKSPACE_BYTES, st));
  w2a::FeatArgs a;
  a.county_start = reinterpret_cast<const long long*>(county_start);
  a.date_days = date_days;
  a.himax_c = himax_c;
  a.alert = alert;
  a.issued = issued_in_advance;
  a.holiday = holiday;
  a.feat = feat;
  a.year = year;
  a.hist = hist;
  a.S = n_counties;
  a.max_rows = max_rows_per_county;
  if (smem > 48 * 1024)
    W2A_CUDA(ctx, cudaFuncSetAttribute(w2a::feature_county_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
  w2a::feature_county_kernel<<<std::min(n_counties, ctx->sm_count * 4), 256, smem, st>>>(a);
  w2a::bspline_stats_kernel<<<1, 32, 0, st>>>(hist, stats);
  const int grid = static_cast<int>(std::min<int64_t>((n_rows + 255) / 256, static_cast<int64_t>(ctx->sm_count) * 8));
  w2a::bspline_apply_kernel<<<grid, 256, 0, st>>>(feat, stats, static_cast<long long>(n_rows));
  ctx->launches += 3;
  W2A_CUDA(ctx, cudaGetLastError());
  return W2A_OK;
}

}  // extern "C"

namespace {
int fill_policy_args(w2a_ctx* ctx, const char* who, int64_t n, const float* obs, const float* w1, const float* b1, const float* w2,
                     float b2, int hidden, int activation, int stochastic, uint64_t seed, uint32_t step, uint8_t* action,
                     float* logit, w2a::PolicyMlpArgs* a) {
  if (n < 0 || n >= (int64_t{1} << 31) - 64 || !obs || !w1 || !b1 || !w2 || !action)
    return fail(ctx, W2A_E_INVALID, "%s: NULL argument or n out of range", who);
  if (hidden < 8 || hidden > w2a::kPolMaxHidden || (hidden & 7))
    return fail(ctx, W2A_E_INVALID, "%s: hidden=%d must be a multiple of 8 in [8, %d]", who, hidden, w2a::kPolMaxHidden);
  if (activation != W2A_ACT_RELU && activation != W2A_ACT_TANH)
    return fail(ctx, W2A_E_INVALID, "%s: unknown activation %d", who, activation);
  a->obs = obs;
  a->w1 = w1;
  a->b1 = b1;
  a->w2 = w2;
  a->b2 = b2;
  a->hidden = hidden;
  a->activation = activation;
  a->stochastic = stochastic != 0;
  a->seed = seed;
  a->step = step;
  a->action = action;
  a->logit = logit;
  a->n = static_cast<int>(n);
  return W2A_OK;
}

// one instance per (hidden size rounded up to 16 / 32 / 64, activation): the tile loops unroll at compile time
template <typename F>
int by_policy_shape(int hidden, int activation, F&& go) {
  const int nt = hidden <= 16 ? 2 : hidden <= 32 ? 4 : 8;
  if (activation == W2A_ACT_RELU) {
    if (nt == 2) return go(std::integral_constant<int, 2>{}, std::integral_constant<int, 0>{});
    if (nt == 4) return go(std::integral_constant<int, 4>{}, std::integral_constant<int, 0>{});
    return go(std::integral_constant<int, 8>{}, std::integral_constant<int, 0>{});
  }
  if (nt == 2) return go(std::integral_constant<int, 2>{}, std::integral_constant<int, 1>{});
  if (nt == 4) return go(std::integral_constant<int, 4>{}, std::integral_constant<int, 1>{});
  return go(std::integral_constant<int, 8>{}, std::integral_constant<int, 1>{});
}

int launch_policy(w2a_ctx* ctx, cudaStream_t st, const w2a::PolicyMlpArgs& a) {
  constexpr int kBlock = 32 * w2a::kPolWarps;
  const size_t smem = sizeof(w2a::PolicyScratch);
  int rc = by_policy_shape(a.hidden, a.activation, [&](auto nt, auto act) -> int {
    auto kernel = w2a::policy_mlp_kernel<decltype(nt)::value, decltype(act)::value>;
    int grid = 0;
    int rc2 = grid_for(ctx, kernel, kBlock, smem, kBlock, a.n, &grid);
    if (rc2) return rc2;
    W2A_CUDA(ctx, launch_kernel(kernel, grid, kBlock, smem, st, ctx->pdl && ctx->pdl_policy, a));
    return W2A_OK;
  });
  if (rc) return rc;
  ctx->launches++;
  W2A_CUDA(ctx, cudaGetLastError());
  return W2A_OK;
}
}  // namespace

extern "C" {

int w2a_policy_mlp(w2a_ctx* ctx, void* stream, int64_t n, const float* obs, const float* w1, const float* b1,
                   const float* w2, float b2, int hidden, int activation, int stochastic, uint64_t seed, uint32_t step,
                   uint8_t* action, float* logit) {
  if (!ctx) return W2A_E_INVALID;
  if (n == 0) return W2A_OK;
  w2a::PolicyMlpArgs a;
  int rc = fill_policy_args(ctx, "w2a_policy_mlp", n, obs, w1, b1, w2, b2, hidden, activation, stochastic, seed, step, action,
                            logit, &a);
  if (rc) return rc;
  W2A_CUDA(ctx, cudaSetDevice(ctx->device));
  return launch_policy(ctx, static_cast<cudaStream_t>(stream), a);
}

int w2a_step_mlp(w2a_ctx* ctx, void* stream, int64_t n, const uint8_t* action, void* state, float* obs, float* reward,
                 uint8_t* terminated, int32_t* remaining, uint8_t* at_budget, int variant, const float* w1, const float* b1,
                 const float* w2, float b2, int hidden, int activation, int stochastic, uint64_t seed, uint32_t step,
                 uint8_t* next_action, float* logit, double* ep_return) {
  int rc = check_ready(ctx, "w2a_step_mlp", n);
  if (rc) return rc;
  if (n == 0) return W2A_OK;
  if (!action || !state || !obs || !reward || !terminated) return fail(ctx, W2A_E_INVALID, "w2a_step_mlp: NULL argument");
  if (next_action == action) return fail(ctx, W2A_E_INVALID, "w2a_step_mlp: next_action must not alias action (alternate two buffers)");
  w2a::PolicyMlpArgs m;
  rc = fill_policy_args(ctx, "w2a_step_mlp", n, obs, w1, b1, w2, b2, hidden, activation, stochastic, seed, step, next_action,
                        logit, &m);
  if (rc) return rc;
  W2A_CUDA(ctx, cudaSetDevice(ctx->device));
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  if ((variant & ~W2A_STEP_LOCAL) != W2A_STEP_TABLE || !ctx->tb.z32_all) {
    // no tables (or another variant asked for): the same result from two launches, and the return sum from a third
    rc = step_planes(ctx, st, n, action, W2A_ACT_U8, plane_static(state), plane_mut(state, n), obs, reward, terminated, remaining,
                     at_budget, variant);
    if (rc) return rc;
    if (ep_return) {
      w2a::add_rewards_kernel<<<static_cast<int>(std::min<int64_t>((n + 255) / 256, ctx->sm_count * 8)), 256, 0, st>>>(
          reward, ep_return, static_cast<int>(n));
      ctx->launches++;
    }
    return launch_policy(ctx, st, m);
  }
  w2a::StepArgs a;
  a.tb = ctx->tb;
  a.st_static = plane_static(state);
  a.st_mut = plane_mut(state, n);
  a.action = action;
  a.obs = obs;
  a.reward = reward;
  a.terminated = terminated;
  a.remaining = remaining;
  a.at_budget = at_budget;
  a.n = static_cast<int>(n);
  a.keep_state = n <= ctx->keep_state_max_envs ? 1 : 0;
  constexpr int kBlock = 32 * w2a::kTableWarps;
  rc = by_policy_shape(hidden, activation, [&](auto nt, auto act) -> int {
    auto kernel = w2a::step_mlp_kernel<0, decltype(nt)::value, decltype(act)::value>;
    int grid = 0;
    int rc2 = grid_for(ctx, kernel, kBlock, w2a::kStepMlpSmemBytes, kBlock, n, &grid);
    if (rc2) return rc2;
    W2A_CUDA(ctx, launch_kernel(kernel, grid, kBlock, w2a::kStepMlpSmemBytes, st, ctx->pdl && ctx->pdl_policy, a, m, ep_return));
    return W2A_OK;
  });
  if (rc) return rc;
  ctx->launches++;
  W2A_CUDA(ctx, cudaGetLastError());
  return W2A_OK;
}

int64_t w2a_launch_count(const w2a_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
