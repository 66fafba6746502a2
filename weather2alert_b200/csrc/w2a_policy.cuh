// w2a_policy.cuh -- a small MLP policy on the 29-slot observation, evaluated on the device between two steps.
//
// the reference's driver loop is `a = policy(obs); env.step(a)` (env.py:273-276 with a policy in
// place of `action_space.sample()`).  With the step kernel at ~35 us per 2^20 envs, a policy built from framework
// kernels (nan_to_num, two matmuls, an activation, a compare) costs ten times the step; this kernel is the whole
//   logit = w2 . act(W1 . nan_to_num(obs) + b1) + b2,   action = logit > 0   (or a Bernoulli(sigmoid(logit)) draw)
// for a batch in one launch, HBM-bound on the 116 observation bytes it reads per env.
//
// The 29 -> H contraction is the one dense GEMM of the path ([n, 32] x [32, H], H <= 64) and runs on the tensor
// cores: `mma.sync.m16n8k16` with float16 inputs and float32 accumulation, a warp per 32 envs (2 m-tiles x H/8 n-tiles
// x 2 k-steps).  The op reads 116 B and does ~4 kFLOP per env (intensity 35 FLOP/B against a ridge of ~210), so the
// tensor pipe idles most of the time by construction -- the point of using it is to get the 1 900 multiply-adds per
// env off the FP32 pipe (56 us per 2^20 envs there), not to approach its peak; a tcgen05/TMEM pipeline would buy
// nothing here.  Inputs and weights are rounded to float16 (11 significant bits, the precision of TF32; observation
// slots stay far below its range): logits agree with a float32 evaluation to ~1e-3 of the summed |x.w| magnitudes;
// an agent that trains on these rollouts should evaluate its policy the same way.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "w2a_kernels.cuh"

namespace w2a {

constexpr int kPolMaxHidden = 64;
constexpr int kPolK = 32;        // 29 observation slots padded to 4 k-steps of 8
constexpr int kPolWStride = 20;  // shared-memory row stride of W1 in half2 words (16 + 4): conflict-free B-fragment reads

struct PolicyMlpArgs {
  const float* obs;   // [n, 29]
  const float* w1;    // [H, 29] row-major (torch.nn.Linear.weight)
  const float* b1;    // [H]
  const float* w2;    // [H]
  float b2;
  int hidden;         // H: multiple of 8, <= 64
  int activation;     // 0 relu, 1 tanh
  int stochastic;     // 0: action = logit > 0; 1: hash(seed, env, step) < sigmoid(logit) * 2^32
  unsigned long long seed;
  unsigned step;
  uint8_t* action;    // [n]
  float* logit;       // [n] or null
  int n;
};

__device__ __forceinline__ unsigned pack_half2(float lo, float hi) {
  unsigned r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));  // first source -> upper half
  return r;
}
// D(16x8) += A(16x16) * B(16x8), float16 inputs (two per register), float32 accumulate.
// A: a0 = [gid][2tig, 2tig+1], a1 = [gid+8][...], a2 = [gid][2tig+8, +9], a3 = [gid+8][2tig+8, +9];
// B: b0 = [2tig, 2tig+1][gid], b1 = [2tig+8, 2tig+9][gid];  C[gid | gid+8][2*tig + {0,1}]  (gid = lane / 4, tig = lane % 4)
__device__ __forceinline__ void mma_f16(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int kPolWarps = 4;
struct __align__(16) PolicyScratch {
  unsigned w1[kPolMaxHidden * kPolWStride];  // half2 words, [h][k / 2], k >= 29 zero
  float b1[kPolMaxHidden], w2[kPolMaxHidden];
  float rows[kPolWarps][2][32 * kObsDim + 4];  // per warp, double buffered: the 32 observation rows of a group, as in global memory
};

__global__ void __launch_bounds__(32 * kPolWarps, 4) policy_mlp_kernel(const PolicyMlpArgs p) {
  extern __shared__ __align__(16) unsigned char pol_smem[];
  PolicyScratch& s = *reinterpret_cast<PolicyScratch*>(pol_smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int H = p.hidden, ntiles = H >> 3;
  for (int i = threadIdx.x; i < kPolMaxHidden * kPolWStride; i += blockDim.x) {
    const int h = i / kPolWStride, k = 2 * (i - h * kPolWStride);
    const float lo = (h < H && k < kObsDim) ? __ldg(p.w1 + h * kObsDim + k) : 0.f;
    const float hi = (h < H && k + 1 < kObsDim) ? __ldg(p.w1 + h * kObsDim + k + 1) : 0.f;
    s.w1[i] = pack_half2(lo, hi);
  }
  for (int i = threadIdx.x; i < kPolMaxHidden; i += blockDim.x) {
    s.b1[i] = i < H ? __ldg(p.b1 + i) : 0.f;
    s.w2[i] = i < H ? __ldg(p.w2 + i) : 0.f;
  }
  pdl_launch_dependents();
  pdl_wait();  // the observations come from the step launch before this one
  __syncthreads();
  const int n_groups = (p.n + 31) >> 5;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const bool aligned = (reinterpret_cast<uintptr_t>(p.obs) & 15) == 0;
  // the rows of group `wg` -> buffer `b`: 32 x 116 contiguous bytes copied as they lie, asynchronously (cp.async, 16 B
  // per lane and round) when the group is full, so that the next group's rows fly while this one is in the tensor cores
  auto fetch_rows = [&](int wg, int b) {
    if (wg >= n_groups) return;
    const int base = wg << 5;
    const int n_valid = min(32, p.n - base);
    const float* src = p.obs + static_cast<size_t>(base) * kObsDim;
    float* dst = s.rows[warp][b];
    if (n_valid == 32 && aligned) {
      for (int i = lane; i < 32 * kObsDim / 4; i += 32)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(dst + 4 * i))),
                     "l"(src + 4 * i)
                     : "memory");
    } else {
      for (int i = lane; i < 32 * kObsDim; i += 32) dst[i] = i < n_valid * kObsDim ? __ldcs(src + i) : 0.f;
    }
  };
  int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int buf = 0;
  fetch_rows(wg, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (; wg < n_groups; wg += warps_total, buf ^= 1) {
    const int base = wg << 5;
    const int n_valid = min(32, p.n - base);
    fetch_rows(wg + warps_total, buf ^ 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");  // this group's rows have landed; the next group's are in flight
    __syncwarp();
    const float* const rows = s.rows[warp][buf];
    // ---- hidden pre-activations: [32, 32] x [32, H] on the tensor cores -----------------------------------
    float acc[2][kPolMaxHidden / 8][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int j = 0; j < kPolMaxHidden / 8; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[m][j][c] = 0.f;
#pragma unroll
    for (int ks = 0; ks < kPolK / 16; ++ks) {
      unsigned a[2][4];
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int r = 16 * m + gid + 8 * (c & 1), k = 16 * ks + 2 * tig + 8 * (c >> 1);
          float v0 = k < kObsDim ? rows[r * kObsDim + k] : 0.f;
          float v1 = k + 1 < kObsDim ? rows[r * kObsDim + k + 1] : 0.f;
          v0 = v0 == v0 ? v0 : 0.f;  // nan_to_num: the significance slot is NaN when there is no alert
          v1 = v1 == v1 ? v1 : 0.f;
          a[m][c] = pack_half2(v0, v1);
        }
#pragma unroll
      for (int j = 0; j < kPolMaxHidden / 8; ++j) {
        if (j < ntiles) {
          const unsigned* wrow = s.w1 + (8 * j + gid) * kPolWStride + 8 * ks + tig;
          const unsigned b0 = wrow[0], b1 = wrow[4];
          mma_f16(acc[0][j], a[0], b0, b1);
          mma_f16(acc[1][j], a[1], b0, b1);
        }
      }
    }
    // ---- activation, output layer: every thread folds its columns, the four threads of a row then add up --------
    float part[2][2] = {{0.f, 0.f}, {0.f, 0.f}};  // [m-tile][row gid | gid + 8]
#pragma unroll
    for (int j = 0; j < kPolMaxHidden / 8; ++j) {
      if (j < ntiles) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int h = 8 * j + 2 * tig + (c & 1);
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            const float pre = acc[m][j][c] + s.b1[h];
            const float hid = p.activation == 0 ? fmaxf(pre, 0.f) : tanhf(pre);
            part[m][c >> 1] = fmaf(hid, s.w2[h], part[m][c >> 1]);
          }
        }
      }
    }
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float v = part[m][r];
        v += __shfl_xor_sync(kFull, v, 1);
        v += __shfl_xor_sync(kFull, v, 2);
        part[m][r] = v + p.b2;
      }
    if (tig == 0) {
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int row = 16 * m + gid + 8 * r;
          const int e = base + row;
          if (row < n_valid) {
            const float z = part[m][r];
            unsigned a;
            if (p.stochastic) {
              const double thr = static_cast<double>(1.f / (1.f + __expf(-z))) * 4294967296.0;
              const unsigned thr_u = thr >= 4294967295.0 ? 0xffffffffu : static_cast<unsigned>(thr);
              a = rollout_hash(p.seed, static_cast<unsigned>(e), p.step) < thr_u ? 1u : 0u;
            } else {
              a = z > 0.f ? 1u : 0u;
            }
            p.action[e] = static_cast<uint8_t>(a);
            if (p.logit) p.logit[e] = z;
          }
        }
    }
    __syncwarp();
  }
}

}  // namespace w2a
