// tilebench.cu -- developer tool (not product code): feasibility of a county-tiled step kernel for
// locality-sorted env sets.  One CTA handles items of up to 256 consecutive envs of one county: the
// county's float64 coefficient block (100 records x 58 doubles) and the day's rows (11 years x 32
// doubles) are bulk-copied (TMA) into shared memory once per item; then one THREAD per env does the
// whole step from shared memory.  Synthetic data, no parity: timing only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tilebench tools/tilebench.cu && /tmp/tilebench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int P = 100, WS = 58, Y = 11, XS = 32, OBS = 29, ITEM = 256;
struct Item { int start, count, col, row0; };  // row0: first of the Y day-rows of (county, t)

__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* s, const void* g, unsigned bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((unsigned)__cvta_generic_to_shared(s)), "l"(g), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned phase) {
  asm volatile("{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* g, const void* s, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"((unsigned)__cvta_generic_to_shared(s)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

struct __align__(128) Smem {
  double w[P * WS];          // 46,400 B
  double x[Y * XS];          //  2,816 B
  float xo[Y * 32];          //  1,408 B  float32 copies of the rows (observation slots)
  float obs[ITEM * OBS];     // 29,696 B
  uint64_t bar;
};

template <bool STORE_OBS, bool DO_EXP>
__global__ void __launch_bounds__(ITEM, 2) tiled_step(const Item* __restrict__ items, int n_items, const double* __restrict__ W,
                                                      const double* __restrict__ X, const uint2* __restrict__ st_static,
                                                      uint2* __restrict__ st_mut, const uint8_t* __restrict__ action,
                                                      float* __restrict__ obs, float* __restrict__ reward, uint8_t* __restrict__ term,
                                                      int* __restrict__ rem_out, uint8_t* __restrict__ atb_out) {
  extern __shared__ __align__(128) unsigned char raw[];
  Smem& sm = *reinterpret_cast<Smem*>(raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { mbar_init(&sm.bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  unsigned phase = 0;
  for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
    const Item im = items[it];
    if (tid == 0) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      mbar_expect(&sm.bar, P * WS * 8 + Y * XS * 8);
      bulk_g2s(sm.w, W + (size_t)im.col * P * WS, P * WS * 8, &sm.bar);
      bulk_g2s(sm.x, X + (size_t)im.row0 * XS, Y * XS * 8, &sm.bar);
    }
    // per-env inputs while the tile lands
    const int e = im.start + tid;
    const bool valid = tid < im.count;
    uint2 sv = make_uint2(0, 0), mv = make_uint2(0, 0);
    int a = 0;
    if (valid) { sv = __ldcs(st_static + e); mv = __ldcs(st_mut + e); a = __ldcs(action + e) != 0; }
    // integer transition (same work as the product kernel)
    int rem = (int)mv.x; const unsigned pk = mv.y; const int t = pk & 0xff, streak = (pk >> 8) & 0xff; unsigned win = (pk >> 16) & 0x3fff;
    const int nd = sv.x >> 24; const int atb = rem == 0; const int act = (a && atb) ? 0 : a; rem -= act;
    win = ((win << 1) | act) & 0x3fff; const int lag = t > 0 ? act : 0; const int cnt = __popc(win); const int done = t >= nd - 1;
    const int t2 = done ? t : t + 1, s2 = done ? streak : (act ? streak + 1 : 0);
    if (valid) __stcs(st_mut + e, make_uint2((unsigned)rem, (unsigned)t2 | (s2 << 8) | (win << 16) | (done << 30)));
    const int p = sv.y % P, yr = (sv.x & 0xffffff) % Y;
    mbar_wait(&sm.bar, phase); phase ^= 1;
    // float32 copies of the Y rows for the observations: once per item
    for (int i = tid; i < Y * 32; i += ITEM) sm.xo[i] = (float)sm.x[i];
    __syncthreads();
    // dot products from shared memory: 28 (b,e) pairs x 24 features + 4 agent terms
    const double2* wr = reinterpret_cast<const double2*>(sm.w + p * WS);
    const double2* xr = reinterpret_cast<const double2*>(sm.x + yr * XS);
    double zb = wr[27].x, ze = wr[27].y;
#pragma unroll
    for (int k = 0; k < 13; ++k) {  // 26 positions: 24 features + sig (x0) + gate
      const double2 xv = xr[k]; const double2 w0 = wr[2 * k], w1 = wr[2 * k + 1];
      zb = fma(xv.x, w0.x, zb); ze = fma(xv.x, w0.y, ze); zb = fma(xv.y, w1.x, zb); ze = fma(xv.y, w1.y, ze);
    }
    { const double2 w0 = wr[26]; zb = fma((double)lag, w0.x, zb); ze = fma((double)streak, w0.y, ze); zb = fma((double)rem, w0.x, zb); }
    float r;
    if (DO_EXP) { const double b = 1.0 / (1.0 + exp(-zb)); const double ef = act ? 1.0 / (1.0 + exp(-ze)) : 0.0; r = (float)(-6.5789 * b * (1.0 - ef)); }
    else r = (float)(zb + ze);
    if (valid) { __stcs(reward + e, r); term[e] = done; __stcs(rem_out + e, rem); atb_out[e] = atb; }
    if (STORE_OBS) {
      float* row = sm.obs + tid * OBS;
      const float* xo = sm.xo + yr * 32;
#pragma unroll
      for (int s = 0; s < 23; ++s) row[s] = xo[s];
      row[23] = (float)lag; row[24] = (float)streak; row[25] = (float)rem; row[26] = xo[23]; row[27] = xo[24]; row[28] = (float)cnt;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      const int nw = min(32, im.count - warp * 32);
      if (lane == 0 && nw == 32) bulk_s2g(obs + (size_t)(im.start + warp * 32) * OBS, sm.obs + warp * 32 * OBS, 32 * OBS * 4);
    }
    __syncthreads();  // the tile buffers are reused by the next item
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

int main() {
  const int S = 746, T = 153, n = S * 1408;  // 1408 = 44 x 32 envs per county keeps every item 16-byte aligned
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  // envs sorted by county (1405 or 1406 per county), then by year; items never straddle a county
  std::vector<Item> items; std::vector<uint2> hs(n), hm(n);
  srand(1);
  int e = 0;
  for (int c = 0; c < S; ++c) {
    const int cnt = 1408;
    for (int s0 = 0; s0 < cnt; s0 += ITEM) items.push_back({e + s0, std::min(ITEM, cnt - s0), c, (c * T + 77) * Y});
    for (int i = 0; i < cnt; ++i, ++e) {
      const int yr = (int)((long long)i * Y / cnt);
      hs[e] = make_uint2((unsigned)(((c * T + 77) * Y + yr)) | (153u << 24), (unsigned)(c * P + rand() % P));
      hm[e] = make_uint2(5u, 77u);
    }
  }
  printf("%d envs, %zu items, %d SMs\n", n, items.size(), prop.multiProcessorCount);
  double *W, *X; Item* dit; uint2 *ss, *smu; uint8_t *act, *term, *atb; float *obs, *rew; int* rem;
  CK(cudaMalloc(&W, (size_t)S * P * WS * 8)); CK(cudaMalloc(&X, (size_t)S * T * Y * XS * 8));
  CK(cudaMemset(W, 0, (size_t)S * P * WS * 8)); CK(cudaMemset(X, 0, (size_t)S * T * Y * XS * 8));
  CK(cudaMalloc(&dit, items.size() * sizeof(Item))); CK(cudaMemcpy(dit, items.data(), items.size() * sizeof(Item), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&ss, n * 8)); CK(cudaMalloc(&smu, n * 8)); CK(cudaMemcpy(ss, hs.data(), n * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(smu, hm.data(), n * 8, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&act, n)); CK(cudaMemset(act, 1, n)); CK(cudaMalloc(&term, n)); CK(cudaMalloc(&atb, n));
  CK(cudaMalloc(&obs, (size_t)n * OBS * 4)); CK(cudaMalloc(&rew, n * 4)); CK(cudaMalloc(&rem, n * 4));
  auto run = [&](auto kern, const char* name) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
    int per_sm = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ITEM, sizeof(Smem)));
    const int grid = prop.multiProcessorCount * per_sm;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) kern<<<grid, ITEM, sizeof(Smem)>>>(dit, (int)items.size(), W, X, ss, smu, act, obs, rew, term, rem, atb);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < 20; ++i) kern<<<grid, ITEM, sizeof(Smem)>>>(dit, (int)items.size(), W, X, ss, smu, act, obs, rew, term, rem, atb);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-34s %d CTAs/SM: %7.1f us per step of %d envs -> %.3e env-steps/s\n", name, per_sm, ms * 50, n, n / (ms * 50e-6));
  };
  run(tiled_step<true, true>, "tiled step (obs + exp)");
  run(tiled_step<true, false>, "tiled step (obs, no exp)");
  run(tiled_step<false, true>, "tiled step (no obs, exp)");
  run(tiled_step<false, false>, "tiled step (no obs, no exp)");
  return 0;
}
