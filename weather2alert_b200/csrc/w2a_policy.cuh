// w2a_policy.cuh -- a small MLP policy on the 29-slot observation, evaluated on the device between two steps
// (policy_mlp_kernel) or inside the step launch itself (step_mlp_kernel, at the end of this file).
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
// nothing here.  Inputs, weights, bias and hidden activations are rounded to float16 (11 significant bits, the
// precision of TF32; values beyond its range are clamped to +-65504) and the second layer runs on the tensor cores as
// well, straight from the first layer's accumulator registers: logits agree with a float32 evaluation to ~1e-3 of the
// summed |x.w| magnitudes; an agent that trains on these rollouts should evaluate its policy the same way.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "w2a_kernels.cuh"

namespace w2a {

constexpr int kPolMaxHidden = 64;
constexpr int kPolK = 32;        // 29 observation slots + the constant-1 input of the bias (k = 29), padded to 2 k-steps of 16
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

// two float32 -> one register of two float16 (first argument in the low half), clamped to the finite float16 range
__device__ __forceinline__ unsigned pack_half2(float lo, float hi) {
  unsigned r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));  // first source -> upper half
  return r;
}
// the same with max(x, 0) applied first: relu, rounding and packing of two hidden units in ONE instruction (F2FP.RELU)
__device__ __forceinline__ unsigned pack_half2_relu(float lo, float hi) {
  unsigned r;
  asm("cvt.rn.satfinite.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float tanh_approx(float x) {  // MUFU.TANH: 2^-11 relative, the precision of the float16 operands
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// nan_to_num on a packed pair: a half that is NaN becomes 0 (the significance slot is NaN when there is no alert)
__device__ __forceinline__ unsigned zero_nan_half2(unsigned v) {
  unsigned m;
  asm("set.eq.u32.f16x2 %0, %1, %1;" : "=r"(m) : "r"(v));  // 0xffff per half that equals itself
  return v & m;
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
constexpr int kPolW1Words = kPolMaxHidden * kPolWStride;  // half2 words of the first layer in shared memory
struct __align__(16) PolicyScratch {
  unsigned w1[kPolW1Words];                    // [h][k / 2]: k < 29 the weights, k = 29 the bias b1[h], k >= 30 zero
  float rows[kPolWarps][2][32 * kObsDim + 4];  // per warp, double buffered: the 32 observation rows of a group, as in global memory
};

// NT = n-tiles of 8 hidden units (H rounded up to 16 / 32 / 64: units beyond H have zero weights and contribute 0),
// ACT = 0 relu, 1 tanh.
// First layer -> shared memory (float16 pairs, b1 as the weight of the constant-1 input in the padding column k = 29);
// second layer -> this lane's B fragments of the [16 x 8] tiles whose column 0 is w2 (the other seven columns are zero).
template <int NT>
__device__ __forceinline__ void mlp_load_weights(const PolicyMlpArgs& p, unsigned* w1s, unsigned (&w2f)[NT / 2][2]) {
  const int H = p.hidden;
  const int gid = (threadIdx.x & 31) >> 2, tig = threadIdx.x & 3;
  for (int i = threadIdx.x; i < 8 * NT * kPolWStride; i += blockDim.x) {
    const int h = i / kPolWStride, k = 2 * (i - h * kPolWStride);
    const float lo = (h < H && k < kObsDim) ? __ldg(p.w1 + h * kObsDim + k) : 0.f;
    float hi = (h < H && k + 1 < kObsDim) ? __ldg(p.w1 + h * kObsDim + k + 1) : 0.f;
    if (h < H && k + 1 == kObsDim) hi = __ldg(p.b1 + h);
    w1s[i] = pack_half2(lo, hi);
  }
#pragma unroll
  for (int u = 0; u < NT / 2; ++u) {
    const int h0 = 16 * u + 2 * tig;
    const float v0 = (gid == 0 && h0 < H) ? __ldg(p.w2 + h0) : 0.f, v1 = (gid == 0 && h0 + 1 < H) ? __ldg(p.w2 + h0 + 1) : 0.f;
    const float v8 = (gid == 0 && h0 + 8 < H) ? __ldg(p.w2 + h0 + 8) : 0.f, v9 = (gid == 0 && h0 + 9 < H) ? __ldg(p.w2 + h0 + 9) : 0.f;
    w2f[u][0] = pack_half2(v0, v1);
    w2f[u][1] = pack_half2(v8, v9);
  }
}

// The policy on one group of 32 envs whose observation rows lie in shared memory as in the observation tensor (row
// stride 29 floats).  Per lane: 32 shared loads + 16 conversions for the A fragments, then per 16 hidden units 8
// tensor-core instructions for the first layer, 4 conversions (activation included) and 2 more tensor-core instructions
// for the second layer: the first layer's accumulator fragment of two neighbouring n-tiles IS the A fragment of a
// 16 x 16 tile, so the hidden activations never leave the registers.
// MT = m-tiles (16 envs each) per pass: 2 shares every first-layer B fragment between both halves of the group; 1 does
// the halves one after the other with half the registers (the step kernel hosting the policy has few to spare).
template <int NT, int ACT, int MT>
__device__ __forceinline__ void mlp_group(const PolicyMlpArgs& p, const float* rows, const unsigned* w1s,
                                          const unsigned (&w2f)[NT / 2][2], int base, int n_valid) {
  const int lane = threadIdx.x & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const unsigned w1_lane = static_cast<unsigned>(__cvta_generic_to_shared(w1s + (lane & 7) * kPolWStride + (lane >> 3) * 4));
#pragma unroll
  for (int m0 = 0; m0 < 2; m0 += MT) {
    // ---- A fragments: [MT m-tiles of 16 envs] x [2 k-steps of 16 slots] ------------------------------------------
    unsigned a[MT][2][4];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float* el = rows + (16 * (m0 + m) + gid + 8 * (c & 1)) * kObsDim + 16 * ks + 2 * tig + 8 * (c >> 1);
          float v0, v1;
          if (ks == 1 && (c >> 1) == 1) {  // slots 24 + 2 tig: (24, 25) (26, 27) (28, bias input 1) (none)
            v0 = tig < 3 ? el[0] : 0.f;
            v1 = tig < 2 ? el[1] : (tig == 2 ? 1.f : 0.f);
          } else {
            v0 = el[0];
            v1 = el[1];
          }
          a[m][ks][c] = zero_nan_half2(pack_half2(v0, v1));
        }
    // ---- 16 hidden units at a time: first layer, activation, second layer -- all in registers --------------------
    float out[MT][4];  // column 0 of C = the logit of rows gid | gid + 8 of the m-tile
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int c = 0; c < 4; ++c) out[m][c] = 0.f;
#pragma unroll
    for (int u = 0; u < NT / 2; ++u) {
      float acc[2][MT][4];  // [n-tile of the pair][m-tile]
#pragma unroll
      for (int jj = 0; jj < 2; ++jj)
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[jj][m][c] = 0.f;
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        // the four B fragments of n-tile 2u + jj (8 hidden units x 32 slots) in one instruction: lane i hands in the
        // address of row i % 8, 16-byte column i / 8 of the tile (80-byte row stride: conflict-free)
        unsigned b[4];
        const unsigned waddr = w1_lane + static_cast<unsigned>((8 * (2 * u + jj)) * kPolWStride * sizeof(unsigned));
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3])
                     : "r"(waddr));
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
          for (int m = 0; m < MT; ++m) mma_f16(acc[jj][m], a[m][ks], b[2 * ks], b[2 * ks + 1]);
      }
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        unsigned hfrag[4];  // C layout of n-tiles (2u, 2u + 1) = A layout of the 16 x 16 tile of hidden units 16u .. 16u + 15
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float x0 = acc[c >> 1][m][2 * (c & 1)], x1 = acc[c >> 1][m][2 * (c & 1) + 1];
          hfrag[c] = ACT == 0 ? pack_half2_relu(x0, x1) : pack_half2(tanh_approx(x0), tanh_approx(x1));
        }
        mma_f16(out[m], hfrag, w2f[u][0], w2f[u][1]);
      }
    }
    if (tig == 0) {
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int row = 16 * (m0 + m) + gid + 8 * r;
          const int e = base + row;
          if (row < n_valid) {
            const float z = out[m][2 * r] + p.b2;
            unsigned act;
            if (p.stochastic) {
              const double thr = static_cast<double>(1.f / (1.f + __expf(-z))) * 4294967296.0;
              const unsigned thr_u = thr >= 4294967295.0 ? 0xffffffffu : static_cast<unsigned>(thr);
              act = rollout_hash(p.seed, static_cast<unsigned>(e), p.step) < thr_u ? 1u : 0u;
            } else {
              act = z > 0.f ? 1u : 0u;
            }
            p.action[e] = static_cast<uint8_t>(act);
            if (p.logit) p.logit[e] = z;
          }
        }
    }
  }
}

template <int NT, int ACT>
__global__ void __launch_bounds__(32 * kPolWarps, 5) policy_mlp_kernel(const PolicyMlpArgs p) {
  extern __shared__ __align__(16) unsigned char pol_smem[];
  PolicyScratch& s = *reinterpret_cast<PolicyScratch*>(pol_smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_launch_dependents();
  pdl_wait();  // the observations come from the step launch before this one (and the weights may come from an optimiser's)
  unsigned w2f[NT / 2][2];
  mlp_load_weights<NT>(p, s.w1, w2f);
  __syncthreads();
  const int n_groups = (p.n + 31) >> 5;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const bool aligned = (reinterpret_cast<uintptr_t>(p.obs) & 15) == 0;
  // Groups are walked from the LAST one down: the step kernel that produced the observations wrote them in ascending
  // order, so the tail of the buffer is what is still in L2 when this kernel starts.
  // The rows of a group -> buffer `b`: 32 x 116 contiguous bytes copied as they lie, asynchronously (cp.async, 16 B
  // per lane and round) when the group is full, so that the next group's rows fly while this one is in the tensor cores
  auto fetch_rows = [&](int it, int b) {
    if (it >= n_groups) return;
    const int base = (n_groups - 1 - it) << 5;
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
    const int base = (n_groups - 1 - wg) << 5;
    fetch_rows(wg + warps_total, buf ^ 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");  // this group's rows have landed; the next group's are in flight
    __syncwarp();
    mlp_group<NT, ACT, 2>(p, s.rows[warp][buf], s.w1, w2f, base, min(32, p.n - base));
    __syncwarp();
  }
}

// ---- step + policy in ONE launch ---------------------------------------------------------------------------------
// `a = policy(obs); obs = env.step(a)` step by step costs the observation tensor twice: 116 B per env written by the
// step and read back by the policy, most of it through HBM (122 MB at 2^20 envs, about what L2 holds).  Here the TABLE
// step's warp evaluates the policy on its 32 fresh rows while they sit in its staging buffer, on their way out: the
// launch takes the actions of step k, writes observation, reward and flags of step k as the plain step does, AND the
// policy's actions for step k + 1 -- the tensor-core work fills issue slots the memory-bound step leaves idle.
#ifndef W2A_STEP_MLP_ROWS
#define W2A_STEP_MLP_ROWS 1  // rounds of row gathers in flight in the fused kernel (B200, 7 CTAs/SM: 1 / 2 / 4 rounds = 52.6 / 53.1 / 56.3 us)
#endif
#ifndef W2A_STEP_MLP_MT
#define W2A_STEP_MLP_MT 1  // m-tiles of the policy per pass inside the step kernel (1: half the registers, occupancy matters more)
#endif
template <int NT, int ACT>
struct StepMlpPost {
  const PolicyMlpArgs& m;
  double* ep_return;  // optional: += reward per env (the return of the rollout so far)
  unsigned* w1s;
  unsigned w2f[NT / 2][2];
  __device__ __forceinline__ void reward(int e, float r) const {
    // a reduction without a return value (RED.ADD.F64, done by L2): a load-add-store here would put a DRAM round trip
    // into every group's critical path
    if (ep_return) atomicAdd(ep_return + e, static_cast<double>(r));
  }
  __device__ __forceinline__ void group(const float* rows, int base, int n_valid) const {
    mlp_group<NT, ACT, W2A_STEP_MLP_MT>(m, rows, w1s, w2f, base, n_valid);
  }
};
// ep_return += reward, for the two-launch form of w2a_step_mlp
__global__ void __launch_bounds__(256) add_rewards_kernel(const float* __restrict__ reward, double* __restrict__ ep_return, int n) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) ep_return[e] += static_cast<double>(reward[e]);
}
constexpr int kStepMlpSmemBytes = kTableSmemBytes + kPolW1Words * static_cast<int>(sizeof(unsigned));

#ifndef W2A_STEP_MLP_MINBLOCKS
#define W2A_STEP_MLP_MINBLOCKS 7  // x 4 warps at 72 registers (B200 sweep, us per iteration at 2^20 envs: 4: 64.1, 5: 59.4, 6: 58.2, 7: 57.8, 8: 58.0)
#endif
template <int ADT, int NT, int ACT>
__global__ void __launch_bounds__(32 * kTableWarps, W2A_STEP_MLP_MINBLOCKS) step_mlp_kernel(const StepArgs p, const PolicyMlpArgs m, double* ep_return) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  StepMlpPost<NT, ACT> post{m, ep_return, reinterpret_cast<unsigned*>(smem_raw + kTableSmemBytes), {}};
  // the weights go to shared memory with every warp of the CTA present (the body retires warps that have no group),
  // and after the wait: an optimiser's launch may have written them just before
  pdl_launch_dependents();
  pdl_wait();
  mlp_load_weights<NT>(m, post.w1s, post.w2f);
  __syncthreads();
  step_table_body<ADT, W2A_STEP_MLP_ROWS, false>(p, post);
}

}  // namespace w2a
