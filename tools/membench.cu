// membench.cu -- developer tool (not product code): what does the B200 memory system give
// for the access pattern of the step kernel?  Per env: 2 random 128 B lines from a large
// table (X, 321 MB), 2 random lines from a small one (W, 19 MB), 8 lanes per env with one
// 128-bit load per lane and line; optionally a 128 B streaming store per env.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/membench tools/membench.cu && /tmp/membench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int DEPTH, bool LOADX, bool LOADW, bool STORE>
__global__ void gather(const float4* __restrict__ X, const float4* __restrict__ W, const uint2* __restrict__ idx,
                       float4* __restrict__ out, float* __restrict__ sink, int n) {
  const int lane = threadIdx.x & 31, q = lane & 7;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const int n_groups = n >> 5;
  float acc = 0.f;
  for (int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; wg < n_groups; wg += warps_total) {
    const int base = wg << 5;
    const uint2 my = __ldcs(idx + base + lane);
    // rounds of 4 envs; DEPTH rounds of loads in flight
#pragma unroll
    for (int r0 = 0; r0 < 8; r0 += DEPTH) {
      float4 xa[DEPTH], xb[DEPTH], wa[DEPTH], wb[DEPTH];
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) {
        const unsigned row = __shfl_sync(0xffffffffu, my.x, r0 + d, 8);
        const unsigned cv = __shfl_sync(0xffffffffu, my.y, r0 + d, 8);
        if (LOADX) { xa[d] = __ldg(X + (size_t)row * 16 + q); xb[d] = __ldg(X + (size_t)row * 16 + 8 + q); }
        if (LOADW) { wa[d] = __ldg(W + (size_t)cv * 16 + q); wb[d] = __ldg(W + (size_t)cv * 16 + 8 + q); }
      }
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) {
        float4 v = make_float4(0, 0, 0, 0);
        if (LOADX) { v.x += xa[d].x + xb[d].y; v.y += xa[d].y + xb[d].x; }
        if (LOADW) { v.z += wa[d].x + wb[d].y; v.w += wa[d].y + wb[d].x; }
        acc += v.x + v.y + v.z + v.w;
        if (STORE) __stcs(out + (size_t)(base + (lane & 24) + r0 + d) * 8 + q, v);
      }
    }
  }
  if (acc == 12345.678f) *sink = acc;
}

template <int DEPTH, bool LX, bool LW, bool ST>
void run(const char* name, const float4* X, const float4* W, const uint2* idx, float4* out, float* sink, int n, int sms) {
  for (int per_sm : {2, 4, 6, 8}) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    auto k = gather<DEPTH, LX, LW, ST>;
    for (int i = 0; i < 3; ++i) k<<<sms * per_sm, 256>>>(X, W, idx, out, sink, n);
    CK(cudaEventRecord(a));
    for (int i = 0; i < 20; ++i) k<<<sms * per_sm, 256>>>(X, W, idx, out, sink, n);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double us = ms * 1000 / 20;
    const double bytes = (double)n * ((LX ? 256 : 0) + (LW ? 256 : 0) + (ST ? 128 : 0) + 8);
    printf("%-28s depth %d ctas/sm %d : %7.1f us  %6.2f TB/s (L2-level bytes)\n", name, DEPTH, per_sm, us, bytes / us / 1e6);
  }
}

int main() {
  const int n = 1 << 20, RX = 746 * 11 * 153, RW = 74600;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  float4 *X, *W, *out; uint2* idx; float* sink;
  CK(cudaMalloc(&X, (size_t)RX * 256)); CK(cudaMalloc(&W, (size_t)RW * 256));
  CK(cudaMalloc(&out, (size_t)n * 128)); CK(cudaMalloc(&idx, (size_t)n * 8)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(X, 0, (size_t)RX * 256)); CK(cudaMemset(W, 0, (size_t)RW * 256));
  for (int mode = 0; mode < 2; ++mode) {
    std::vector<uint2> h(n);
    srand(1);
    for (int i = 0; i < n; ++i) {
      // mode 0: the bench's pattern -- all envs on the same day t, county = i mod 746, random year / sample
      // mode 1: sorted by (county, sample): env i -> county i / 1406, sample ascending inside the county
      int county, p, y = rand() % 11;
      if (mode == 0) { county = i % 746; p = rand() % 100; }
      else { county = i / 1406; if (county > 745) county = 745; p = (i % 1406) * 100 / 1406; }
      h[i].x = (county * 11 + y) * 153 + 77;
      h[i].y = county * 100 + p;
    }
    CK(cudaMemcpy(idx, h.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
    printf("== %s, %d SMs\n", mode == 0 ? "interleaved counties (bench pattern)" : "sorted by (county, sample)", prop.multiProcessorCount);
    const int sms = prop.multiProcessorCount;
    run<1, true, true, true>("X+W gather + 128B store", X, W, idx, out, sink, n, sms);
    run<2, true, true, true>("X+W gather + 128B store", X, W, idx, out, sink, n, sms);
    run<4, true, true, true>("X+W gather + 128B store", X, W, idx, out, sink, n, sms);
    run<2, true, true, false>("X+W gather only", X, W, idx, out, sink, n, sms);
    run<2, true, false, true>("X gather + store", X, W, idx, out, sink, n, sms);
    run<2, false, true, true>("W gather + store", X, W, idx, out, sink, n, sms);
    run<2, false, false, true>("store only", X, W, idx, out, sink, n, sms);
  }
  return 0;
}
