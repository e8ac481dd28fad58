// ppo_kernels.cu -- the three HBM-bound reductions / scans that sit right after a rollout:
//
//   K4 moments_kernel    n, sum x, sum x^2 of the advantage buffer (one read of [T*n] floats)
//      normalize_kernel  (x - mean) / (std + eps), unbiased std like torch.std
//                        (Actor_Critic_PPO.ipynb c21:L105, Policy_Gradients.ipynb c26:L33-36)
//   N3 gae_kernel        compute_gae backward scan, one thread per env, coalesced over n
//                        (Actor_Critic_PPO.ipynb c15:L49-53)
//
// Streaming float work: float4 loads, grid = a whole number of waves over the 148 SMs, double
// accumulation only in the per-thread / per-warp partials.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/drone_b200.h"
#include "host_guard.h"

namespace dd {

constexpr int kRedBlock = 256;
constexpr int kSMs = 148;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kRedBlock) moments_kernel(const float* __restrict__ x, int64_t n, double* out)
{
    __shared__ double s_s[kRedBlock / 32], s_q[kRedBlock / 32];
    double s = 0.0, q = 0.0;
    const int64_t tid = (int64_t)blockIdx.x * kRedBlock + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * kRedBlock;
    // leading scalars until x is 16-byte aligned, then float4, then the tail
    int64_t head = (int64_t)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(x) & 15u)) & 15u) / 4;
    if (head > n) head = n;
    const int64_t n4 = (n - head) / 4;
    const float4* x4 = reinterpret_cast<const float4*>(x + head);
    for (int64_t j = tid; j < n4; j += nthreads) {
        const float4 v = __ldg(x4 + j);
        // partial sums of 4 in float would lose bits; go to double per element
        s += (double)v.x + (double)v.y + (double)v.z + (double)v.w;
        q += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    if (tid < head) { const double v = x[tid]; s += v; q += v * v; }
    const int64_t tail0 = head + 4 * n4;
    if (tid < n - tail0) { const double v = x[tail0 + tid]; s += v; q += v * v; }

    s = warp_sum_d(s); q = warp_sum_d(q);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { s_s[w] = s; s_q[w] = q; }
    __syncthreads();
    if (w == 0) {
        s = lane < kRedBlock / 32 ? s_s[lane] : 0.0;
        q = lane < kRedBlock / 32 ? s_q[lane] : 0.0;
        s = warp_sum_d(s); q = warp_sum_d(q);
        if (lane == 0) {
            atomicAdd(out + 1, s);
            atomicAdd(out + 2, q);
            if (blockIdx.x == 0) atomicAdd(out + 0, (double)n);
        }
    }
}

__global__ void __launch_bounds__(kRedBlock) normalize_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                               const double* __restrict__ m, double eps, int64_t n)
{
    const double cnt = m[0], s = m[1], q = m[2];
    const double mean = cnt > 0 ? s / cnt : 0.0;
    double var = cnt > 1 ? (q - s * mean) / (cnt - 1.0) : 0.0;     // unbiased, torch.std default
    var = var > 0 ? var : 0.0;
    const float fmean = (float)mean;
    const float inv = (float)(1.0 / (sqrt(var) + eps));
    const int64_t tid = (int64_t)blockIdx.x * kRedBlock + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * kRedBlock;
    const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15u) == 0;
    if (vec) {
        const int64_t n4 = n / 4;
        const float4* x4 = reinterpret_cast<const float4*>(x);
        float4* y4 = reinterpret_cast<float4*>(y);
        for (int64_t j = tid; j < n4; j += nthreads) {
            float4 v = __ldg(x4 + j);
            v.x = (v.x - fmean) * inv; v.y = (v.y - fmean) * inv; v.z = (v.z - fmean) * inv; v.w = (v.w - fmean) * inv;
            y4[j] = v;
        }
        if (tid < n - 4 * n4) y[4 * n4 + tid] = (x[4 * n4 + tid] - fmean) * inv;
    } else {
        for (int64_t j = tid; j < n; j += nthreads) y[j] = (x[j] - fmean) * inv;
    }
}

// delta_t = r_t + gamma * V_{t+1} * (1 - done_t) - V_t ;  A_t = delta_t + gamma * lambda * (1 - done_t) * A_{t+1}
// fp32 throughout, in the operation order of the torch original (python scalars gamma, gamma*lambda
// are combined in double and rounded to fp32 when they meet the tensors).  No FMA contraction, so
// the result is bit-identical to the eager torch loop.
// The scan is a dependent chain only through `gae` (4 dependent fp32 operations per step); the loads are not on
// it.  A thread-per-env scan has little parallelism at PPO sizes (65,536 envs = 443 threads per SM), so the
// loop works in chunks of kGaeUnroll steps whose 3 x kGaeUnroll loads are all issued before the first is used:
// the launch is bound by DRAM latency / kGaeUnroll instead of DRAM latency per step.
// Chunks are double-buffered in registers: the loads of chunk c+1 are issued BEFORE chunk c is scanned, so loads
// stay in flight during the dependent arithmetic and the stores (round 1 issued them only after the scan of the
// previous chunk: 4.4 TB/s; with the prefetch the kernel is bound by HBM, not by its latency).
#ifndef DD_GAE_BLOCK
#define DD_GAE_BLOCK 64
#endif
#ifndef DD_GAE_UNROLL
#define DD_GAE_UNROLL 10
#endif
constexpr int kGaeBlock = DD_GAE_BLOCK;
constexpr int kGaeUnroll = DD_GAE_UNROLL;

struct GaeChunk {
    float r[kGaeUnroll], v[kGaeUnroll];
    uint8_t d[kGaeUnroll];
    // steps t1-1 ... t1-kGaeUnroll (clamped at 0: the surplus loads of the last chunk re-read row 0)
    __device__ __forceinline__ void load(const float* __restrict__ rew, const float* __restrict__ val,
                                         const uint8_t* __restrict__ done, int32_t t1, int64_t n, int64_t i) {
#pragma unroll
        for (int j = 0; j < kGaeUnroll; ++j) {
            const int32_t t = t1 - 1 - j;
            const int64_t o = (int64_t)(t >= 0 ? t : 0) * n + i;
            r[j] = __ldg(rew + o); v[j] = __ldg(val + o); d[j] = __ldg(done + o);
        }
    }
    __device__ __forceinline__ void scan(float* __restrict__ adv, float* __restrict__ ret, float g, float gl, int32_t t1,
                                         int64_t n, int64_t i, float& gae, float& v_next) const {
#pragma unroll
        for (int j = 0; j < kGaeUnroll; ++j) {
            const int32_t t = t1 - 1 - j;
            if (t >= 0) {
                const int64_t o = (int64_t)t * n + i;
                const float mask = d[j] ? 0.0f : 1.0f;
                const float delta = __fsub_rn(__fadd_rn(r[j], __fmul_rn(__fmul_rn(g, v_next), mask)), v[j]);
                gae = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, mask), gae));
                adv[o] = gae;
                if (ret) ret[o] = __fadd_rn(gae, v[j]);
                v_next = v[j];
            }
        }
    }
};

__global__ void __launch_bounds__(kGaeBlock) gae_kernel(const float* __restrict__ rew, const float* __restrict__ val,
                                                        const uint8_t* __restrict__ done, float* __restrict__ adv,
                                                        float* __restrict__ ret, float g, float gl, int32_t T, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * kGaeBlock + threadIdx.x;
    if (i >= n) return;
    float gae = 0.0f;
    float v_next = __ldg(val + (int64_t)T * n + i);
    GaeChunk a, b;
    a.load(rew, val, done, T, n, i);
    for (int32_t t1 = T; t1 > 0; t1 -= 2 * kGaeUnroll) {       // two chunks per trip: a = [t1-U, t1), b = [t1-2U, t1-U)
        if (t1 - kGaeUnroll > 0) b.load(rew, val, done, t1 - kGaeUnroll, n, i);
        a.scan(adv, ret, g, gl, t1, n, i, gae, v_next);
        if (t1 - kGaeUnroll <= 0) break;
        if (t1 - 2 * kGaeUnroll > 0) a.load(rew, val, done, t1 - 2 * kGaeUnroll, n, i);
        b.scan(adv, ret, g, gl, t1 - kGaeUnroll, n, i, gae, v_next);
    }
}

// compute_returns (Policy_Gradients.ipynb: G = r + gamma * G over reversed(rewards), python floats = float64),
// batched: G restarts after a step that ended an episode.  Accumulated in double like the original, written as
// fp32 (the notebook builds a float32 tensor from the list).  Same chunked load scheme as gae_kernel.
__global__ void __launch_bounds__(kGaeBlock) returns_kernel(const float* __restrict__ rew, const uint8_t* __restrict__ done,
                                                            float* __restrict__ out, double gamma, int32_t T, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * kGaeBlock + threadIdx.x;
    if (i >= n) return;
    double G = 0.0;
    for (int32_t t1 = T; t1 > 0; t1 -= kGaeUnroll) {
        float r[kGaeUnroll];
        uint8_t d[kGaeUnroll];
#pragma unroll
        for (int j = 0; j < kGaeUnroll; ++j) {
            const int32_t t = t1 - 1 - j;
            const int64_t o = (int64_t)(t >= 0 ? t : 0) * n + i;
            r[j] = __ldg(rew + o); d[j] = done ? __ldg(done + o) : (uint8_t)0;
        }
#pragma unroll
        for (int j = 0; j < kGaeUnroll; ++j) {
            const int32_t t = t1 - 1 - j;
            if (t >= 0) {
                if (d[j]) G = 0.0;                              // step t ended its episode: nothing flows back into it
                G = __dadd_rn((double)r[j], __dmul_rn(gamma, G));
                out[(int64_t)t * n + i] = (float)G;
            }
        }
    }
}

static inline int wave_grid(int64_t work_items, int per_block, int max_waves_ctas)
{
    int64_t g = (work_items + per_block - 1) / per_block;
    if (g > max_waves_ctas) g = max_waves_ctas;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace dd

extern "C" {

int dd_moments(const float* x, int64_t n, double* out, void* stream)
{
    if (!x || !out) return DD_E_NULL;
    if (n < 0) return DD_E_RANGE;
    if ((reinterpret_cast<uintptr_t>(x) & 3u) || (reinterpret_cast<uintptr_t>(out) & 7u)) return DD_E_ALIGN;
    if (n == 0) return 0;
    const int grid = dd::wave_grid(n / 4 + 1, dd::kRedBlock * 4, dd::kSMs * 8);
    dd::DeviceGuard guard((cudaStream_t)stream, x);
    if (guard.err != cudaSuccess) return (int)guard.err;
    dd::moments_kernel<<<grid, dd::kRedBlock, 0, (cudaStream_t)stream>>>(x, n, out);
    return (int)cudaGetLastError();
}

int dd_normalize(const float* x, float* y, const double* moments, double eps, int64_t n, void* stream)
{
    if (!x || !y || !moments) return DD_E_NULL;
    if (n < 0) return DD_E_RANGE;
    if (n == 0) return 0;
    const int grid = dd::wave_grid(n / 4 + 1, dd::kRedBlock * 2, dd::kSMs * 8);
    dd::DeviceGuard guard((cudaStream_t)stream, x);
    if (guard.err != cudaSuccess) return (int)guard.err;
    dd::normalize_kernel<<<grid, dd::kRedBlock, 0, (cudaStream_t)stream>>>(x, y, moments, eps, n);
    return (int)cudaGetLastError();
}

int dd_gae(const float* rewards_tn, const float* values_t1n, const uint8_t* dones_tn, float* adv_tn,
           float* returns_tn, double gamma, double lambda, int32_t T, int64_t n, void* stream)
{
    if (!rewards_tn || !values_t1n || !dones_tn || !adv_tn) return DD_E_NULL;
    if (n < 0 || T < 0) return DD_E_RANGE;
    if (n == 0 || T == 0) return 0;
    const int grid = (int)((n + dd::kGaeBlock - 1) / dd::kGaeBlock);
    dd::DeviceGuard guard((cudaStream_t)stream, rewards_tn);
    if (guard.err != cudaSuccess) return (int)guard.err;
    dd::gae_kernel<<<grid, dd::kGaeBlock, 0, (cudaStream_t)stream>>>(rewards_tn, values_t1n, dones_tn, adv_tn, returns_tn,
                                                                       (float)gamma, (float)(gamma * lambda), T, n);
    return (int)cudaGetLastError();
}

int dd_discounted_returns(const float* rewards_tn, const uint8_t* dones_tn, float* returns_tn, double gamma,
                          int32_t T, int64_t n, void* stream)
{
    if (!rewards_tn || !returns_tn) return DD_E_NULL;
    if (n < 0 || T < 0) return DD_E_RANGE;
    if (n == 0 || T == 0) return 0;
    const int grid = (int)((n + dd::kGaeBlock - 1) / dd::kGaeBlock);
    dd::DeviceGuard guard((cudaStream_t)stream, rewards_tn);
    if (guard.err != cudaSuccess) return (int)guard.err;
    dd::returns_kernel<<<grid, dd::kGaeBlock, 0, (cudaStream_t)stream>>>(rewards_tn, dones_tn, returns_tn, gamma, T, n);
    return (int)cudaGetLastError();
}

}  // extern "C"
