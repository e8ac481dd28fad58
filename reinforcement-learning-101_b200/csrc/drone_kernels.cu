// drone_kernels.cu -- sm_100a kernels of the batched delivery_drone simulator + their C ABI.
//
//   K1 step_kernel     one DroneGame.step() (+ get_state, + same-step reset) per env per launch:
//                      one HBM round trip (SoA state in, state + reward + flags + obs out)
//   K2 reset_kernel    DroneGame.reset() for all / masked envs (Philox spawn)
//   K3 episode stats   warp ballot / shuffle -> per-warp partials -> one RED per block into one of
//                      DD_STATS_SLOTS copies (fused into K1 and the rollout kernel)
//   rollout_kernel     T steps per launch with the env state in registers
//
// HBM-bound byte/float work: no tensor cores here.  What matters is 16-byte coalesced loads and
// stores, the observation tile staged through shared memory and written with one TMA bulk store
// per CTA, and a grid that is a whole number of waves over 148 SMs.
#include <cuda_runtime.h>
#include <stdint.h>
#include "drone_core.cuh"

namespace dd {

constexpr int kBlock = 256;            // threads per CTA == envs per CTA per tile
constexpr int kWarps = kBlock / 32;
constexpr int kMaxObsStride = 16;

template <typename R>
struct KArgs {
    R *pos_vel, *att_fuel, *platform;
    int32_t* steps;
    uint32_t* episode;
    uint8_t* flags;
    const uint8_t* actions;
    R *obs, *reward, *final_obs;
    uint8_t* done_flags;
    unsigned long long* stats;
    int64_t n;
    uint64_t seed, env_id_base;
    int32_t max_steps, obs_stride;
    int32_t rand_drone, rand_platform;
    Consts<R> k;
};

// ---- 16-byte vector access ---------------------------------------------------------------
__device__ __forceinline__ void load4(const float* p, int64_t i, float& a, float& b, float& c, float& d) {
    const float4 v = reinterpret_cast<const float4*>(p)[i];
    a = v.x; b = v.y; c = v.z; d = v.w;
}
__device__ __forceinline__ void load4(const double* p, int64_t i, double& a, double& b, double& c, double& d) {
    const double2 u = reinterpret_cast<const double2*>(p)[2 * i], v = reinterpret_cast<const double2*>(p)[2 * i + 1];
    a = u.x; b = u.y; c = v.x; d = v.y;
}
__device__ __forceinline__ void store4(float* p, int64_t i, float a, float b, float c, float d) {
    reinterpret_cast<float4*>(p)[i] = make_float4(a, b, c, d);
}
__device__ __forceinline__ void store4(double* p, int64_t i, double a, double b, double c, double d) {
    reinterpret_cast<double2*>(p)[2 * i] = make_double2(a, b);
    reinterpret_cast<double2*>(p)[2 * i + 1] = make_double2(c, d);
}
__device__ __forceinline__ void load2(const float* p, int64_t i, float& a, float& b) {
    const float2 v = reinterpret_cast<const float2*>(p)[i]; a = v.x; b = v.y;
}
__device__ __forceinline__ void load2(const double* p, int64_t i, double& a, double& b) {
    const double2 v = reinterpret_cast<const double2*>(p)[i]; a = v.x; b = v.y;
}
__device__ __forceinline__ void store2(float* p, int64_t i, float a, float b) {
    reinterpret_cast<float2*>(p)[i] = make_float2(a, b);
}
__device__ __forceinline__ void store2(double* p, int64_t i, double a, double b) {
    reinterpret_cast<double2*>(p)[i] = make_double2(a, b);
}

template <typename R>
__device__ __forceinline__ void load_env(const KArgs<R>& a, int64_t i, Env<R>& e) {
    load4(a.pos_vel, i, e.x, e.y, e.vx, e.vy);
    load4(a.att_fuel, i, e.angle, e.angvel, e.fuel, e.ret);
    load2(a.platform, i, e.px, e.py);
    e.steps = a.steps[i];
}
template <typename R>
__device__ __forceinline__ void store_env(const KArgs<R>& a, int64_t i, const Env<R>& e) {
    store4(a.pos_vel, i, e.x, e.y, e.vx, e.vy);
    store4(a.att_fuel, i, e.angle, e.angvel, e.fuel, e.ret);
    a.steps[i] = e.steps;
}

// ---- K3: episode statistics ------------------------------------------------------------------
// Words: 0 episodes, 1 landed, 2 crashed, 3 truncated, 4 sum_return (int64, 2^-20), 5 sum_length,
// 6 env_steps.  Integer accumulation makes the totals independent of summation order, hence of
// grid shape and GPU count.
struct StatAcc {
    uint32_t episodes = 0, landed = 0, crashed = 0, truncated = 0, env_steps = 0;
    long long ret_fx = 0;
    long long length = 0;
    __device__ __forceinline__ void on_done(uint32_t f, double ret, int32_t steps) {
        episodes += 1;
        landed += (f & DD_LANDED) ? 1u : 0u;
        crashed += (f & DD_CRASHED) ? 1u : 0u;
        truncated += (f & DD_TRUNCATED) ? 1u : 0u;
        ret_fx += __double2ll_rn(ret * DD_RETURN_FIXED_SCALE);
        length += steps;
    }
};

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-level flush: every warp reduces its lanes with shuffles, lane 0 parks the partials in
// shared memory, and after the barrier 7 threads add the non-zero block totals to this block's
// slot with one RED each.
__device__ __forceinline__ void stats_warp_park(const StatAcc& s, unsigned long long (*part)[DD_STATS_WORDS]) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned any = __ballot_sync(0xffffffffu, (s.episodes | s.env_steps) != 0u);
    unsigned long long v[7] = {0, 0, 0, 0, 0, 0, 0};
    if (any) {                                             // warp-uniform
        v[0] = __reduce_add_sync(0xffffffffu, s.episodes);
        v[6] = __reduce_add_sync(0xffffffffu, s.env_steps);
        if (v[0]) {
            v[1] = __reduce_add_sync(0xffffffffu, s.landed);
            v[2] = __reduce_add_sync(0xffffffffu, s.crashed);
            v[3] = __reduce_add_sync(0xffffffffu, s.truncated);
            v[4] = (unsigned long long)warp_sum_ll(s.ret_fx);
            v[5] = (unsigned long long)warp_sum_ll(s.length);
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 7; ++j) part[w][j] = v[j];
    }
}
__device__ __forceinline__ void stats_block_flush(unsigned long long (*part)[DD_STATS_WORDS],
                                                  unsigned long long* stats) {
    // call after __syncthreads()
    if (threadIdx.x < 7) {
        unsigned long long t = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += part[w][threadIdx.x];
        if (t) atomicAdd(stats + (blockIdx.x % DD_STATS_SLOTS) * DD_STATS_WORDS + threadIdx.x, t);
    }
}

// ---- observation tile: shared memory -> global ------------------------------------------------
// A CTA's tile is contiguous in global memory (rows of obs_stride elements, 256 rows).  Full,
// 16-byte aligned tiles leave with ONE TMA bulk store (cp.async.bulk, SASS UBLKCP) issued by one
// thread; ragged tiles fall back to a coalesced copy loop.
template <typename R>
__device__ __forceinline__ void obs_tile_store(const R* s_obs, R* g_tile, int rows, int stride) {
    const uint32_t bytes = (uint32_t)(rows * stride * sizeof(R));
    const bool bulk = (rows == kBlock) && ((reinterpret_cast<uintptr_t>(g_tile) & 15u) == 0) && ((bytes & 15u) == 0);
    if (bulk) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // my smem writes -> async proxy
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(s_obs);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(g_tile), "r"(saddr), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem must outlive the read
        }
    } else {
        __syncthreads();
        const int total = rows * stride;
        for (int j = threadIdx.x; j < total; j += kBlock) g_tile[j] = s_obs[j];
    }
}

// =================================================================================================
// K1: one step per env per launch.
// =================================================================================================
template <typename R, bool AUTO, bool OBS>
__global__ void __launch_bounds__(kBlock) step_kernel(const __grid_constant__ KArgs<R> a)
{
    __shared__ unsigned long long s_part[kWarps][DD_STATS_WORDS];
    __shared__ __align__(128) R s_obs[OBS ? kBlock * kMaxObsStride : 1];

    const int64_t tile0 = (int64_t)blockIdx.x * kBlock;
    const int64_t i = tile0 + threadIdx.x;
    const bool live = i < a.n;

    StatAcc st;
    if (live) {
        Env<R> e;
        load_env(a, i, e);
        const uint32_t act = a.actions[i];
        uint32_t pflags = AUTO ? 0u : (uint32_t)a.flags[i];   // persistent flags are always 0 under auto-reset
        uint32_t oflags = pflags;                              // what this step reports
        R reward = (R)0, speed, dist;
        bool stepped = false;

        if (!(act & DD_ACT_SKIP) && !(pflags & DD_DONE)) {
            stepped = true;
            st.env_steps = 1;
            uint32_t f = step_core(e, act, a.k, reward, speed, dist);
            if (!f && a.max_steps > 0 && e.steps >= a.max_steps) f = DD_DONE | DD_TRUNCATED;
            oflags = f;
            if (f) {
                st.on_done(f, (double)e.ret, e.steps);
                if (a.final_obs) {                             // terminal observation (rare, scattered)
                    R* fo = a.final_obs + i * a.obs_stride;
                    write_obs(e, f, speed, dist, a.k, [&](int j, R v) { fo[j] = v; });
                }
                if (AUTO) {                                    // same-step reset
                    const uint32_t ep = a.episode[i];
                    spawn(e, a.k, a.seed, a.env_id_base + (uint64_t)i, ep, a.rand_drone != 0, a.rand_platform != 0);
                    a.episode[i] = ep + 1;
                    store2(a.platform, i, e.px, e.py);
                    f = 0;
                    speed = (R)0;
                    if (OBS) { R s_; speed_dist(e, s_, dist); }
                }
            }
            pflags = f;
        } else if (OBS) {
            speed_dist(e, speed, dist);                        // frozen / skipped env: obs of the state as is
        }

        if (stepped) {
            store_env(a, i, e);
            if (!AUTO) a.flags[i] = (uint8_t)pflags;
        }
        if (a.reward) a.reward[i] = reward;
        if (a.done_flags) a.done_flags[i] = (uint8_t)oflags;
        if (OBS) {
            R* row = s_obs + threadIdx.x * a.obs_stride;
            write_obs(e, pflags, speed, dist, a.k, [&](int j, R v) { row[j] = v; });
            if (a.obs_stride > DD_OBS_DIM) row[DD_OBS_DIM] = (R)e.steps;     // 16th key: 'steps'
        }
    }

    if (a.stats) stats_warp_park(st, s_part);
    if (OBS) {
        const int64_t rem = a.n - tile0;
        const int rows = rem < kBlock ? (int)rem : kBlock;
        obs_tile_store(s_obs, a.obs + tile0 * a.obs_stride, rows, a.obs_stride);   // contains the barrier
    } else {
        __syncthreads();
    }
    if (a.stats) stats_block_flush(s_part, a.stats);
}

// =================================================================================================
// K2: reset.
// =================================================================================================
template <typename R>
__global__ void __launch_bounds__(kBlock) reset_kernel(const __grid_constant__ KArgs<R> a, const uint8_t* mask)
{
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= a.n) return;
    if (mask && !mask[i]) return;
    Env<R> e;
    const uint32_t ep = a.episode[i];
    spawn(e, a.k, a.seed, a.env_id_base + (uint64_t)i, ep, a.rand_drone != 0, a.rand_platform != 0);
    a.episode[i] = ep + 1;                                     // game_engine.py:91
    store_env(a, i, e);
    store2(a.platform, i, e.px, e.py);
    a.flags[i] = 0;
    if (a.obs) {
        R speed, dist;
        speed_dist(e, speed, dist);
        R* row = a.obs + i * a.obs_stride;
        write_obs(e, 0u, speed, dist, a.k, [&](int j, R v) { row[j] = v; });
        if (a.obs_stride > DD_OBS_DIM) row[DD_OBS_DIM] = (R)0;
    }
}

// =================================================================================================
// Rollout: T steps per launch, env state in registers, one state round trip per launch.
// =================================================================================================
template <typename R>
struct RArgs {
    KArgs<R> a;
    const uint8_t* actions_tn;
    R* reward_tn;
    uint8_t* done_tn;
    R* obs_tn;
    uint32_t t0;
    int32_t T, policy, auto_reset;
};

template <typename R, bool OBS>
__global__ void __launch_bounds__(kBlock) rollout_kernel(const __grid_constant__ RArgs<R> ra)
{
    __shared__ unsigned long long s_part[kWarps][DD_STATS_WORDS];
    __shared__ __align__(128) R s_obs[OBS ? kBlock * kMaxObsStride : 1];
    const KArgs<R>& a = ra.a;
    const int64_t tile0 = (int64_t)blockIdx.x * kBlock;
    const int64_t i = tile0 + threadIdx.x;
    const bool live = i < a.n;
    const int64_t rem = a.n - tile0;
    const int rows = rem < kBlock ? (int)rem : kBlock;

    Env<R> e = {};
    uint32_t pflags = 0, ep = 0;
    bool platform_dirty = false;
    if (live) {
        load_env(a, i, e);
        pflags = a.flags[i];
        ep = a.episode[i];
    }
    const uint64_t gid = a.env_id_base + (uint64_t)i;
    StatAcc st;
    U4 blk = {0, 0, 0, 0};
    uint32_t blk_id = 0xffffffffu;

    for (int32_t t = 0; t < ra.T; ++t) {
        uint32_t oflags = pflags;
        R reward = (R)0, speed = (R)0, dist = (R)0;
        if (live) {
            uint32_t act;
            if (ra.policy == DD_POLICY_TRACE) {
                act = ra.actions_tn[(int64_t)t * a.n + i];
            } else if (ra.policy == DD_POLICY_RANDOM) {
                const uint32_t tt = ra.t0 + (uint32_t)t;
                if ((tt >> 5) != blk_id) { blk = action_block(a.seed, gid, tt); blk_id = tt >> 5; }
                act = action_from_block(blk, tt);
            } else {
                act = (e.vy > (R)1.5) ? DD_ACT_MAIN : 0u;
            }
            if (!(act & DD_ACT_SKIP) && !(pflags & DD_DONE)) {
                st.env_steps += 1;
                uint32_t f = step_core(e, act, a.k, reward, speed, dist);
                if (!f && a.max_steps > 0 && e.steps >= a.max_steps) f = DD_DONE | DD_TRUNCATED;
                oflags = f;
                if (f) {
                    st.on_done(f, (double)e.ret, e.steps);
                    if (ra.auto_reset) {
                        spawn(e, a.k, a.seed, gid, ep, a.rand_drone != 0, a.rand_platform != 0);
                        ep += 1;
                        platform_dirty = true;
                        f = 0;
                        if (OBS) speed_dist(e, speed, dist);
                    }
                }
                pflags = f;
            } else if (OBS) {
                speed_dist(e, speed, dist);
            }
            if (ra.reward_tn) ra.reward_tn[(int64_t)t * a.n + i] = reward;
            if (ra.done_tn) ra.done_tn[(int64_t)t * a.n + i] = (uint8_t)oflags;
            if (OBS) {
                R* row = s_obs + threadIdx.x * a.obs_stride;
                write_obs(e, pflags, speed, dist, a.k, [&](int j, R v) { row[j] = v; });
                if (a.obs_stride > DD_OBS_DIM) row[DD_OBS_DIM] = (R)e.steps;
            }
        }
        if (OBS) {
            obs_tile_store(s_obs, ra.obs_tn + ((int64_t)t * a.n + tile0) * a.obs_stride, rows, a.obs_stride);
            __syncthreads();                                   // tile is reused next step
        }
    }

    if (live) {
        store_env(a, i, e);
        a.flags[i] = (uint8_t)pflags;
        a.episode[i] = ep;
        if (platform_dirty) store2(a.platform, i, e.px, e.py);
    }
    if (a.stats) {
        stats_warp_park(st, s_part);
        __syncthreads();
        stats_block_flush(s_part, a.stats);
    }
}

// ---- small utility kernels ----------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) fill_random_actions_kernel(uint8_t* out, uint64_t seed, uint64_t env_id_base,
                                                                      uint32_t t0, int32_t T, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    U4 blk = {0, 0, 0, 0};
    uint32_t blk_id = 0xffffffffu;
    for (int32_t t = 0; t < T; ++t) {
        const uint32_t tt = t0 + (uint32_t)t;
        if ((tt >> 5) != blk_id) { blk = action_block(seed, env_id_base + (uint64_t)i, tt); blk_id = tt >> 5; }
        out[(int64_t)t * n + i] = (uint8_t)action_from_block(blk, tt);
    }
}

__global__ void __launch_bounds__(kBlock) pack_actions_kernel(const uint8_t* a3, uint8_t* packed, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const uint8_t m = a3[3 * i], l = a3[3 * i + 1], r = a3[3 * i + 2];
    packed[i] = (uint8_t)((m ? DD_ACT_MAIN : 0u) | (l ? DD_ACT_LEFT : 0u) | (r ? DD_ACT_RIGHT : 0u));
}

__global__ void stats_collapse_kernel(const unsigned long long* stats, unsigned long long* out)
{
    const int w = threadIdx.x;
    if (w >= DD_STATS_WORDS) return;
    unsigned long long t = 0;
    for (int s = 0; s < DD_STATS_SLOTS; ++s) t += stats[s * DD_STATS_WORDS + w];
    out[w] = t;
}

// =================================================================================================
// host side of the ABI
// =================================================================================================
static inline int grid_for(int64_t n) { return (int)((n + kBlock - 1) / kBlock); }

template <typename R>
static int fill_args(KArgs<R>& a, const DDState* s, const DDParams* p, const DDEnvConfig* c, int64_t n)
{
    if (!s || !p || !c) return DD_E_NULL;
    if (n < 0 || n > ((int64_t)1 << 40)) return DD_E_RANGE;
    if (n > 0 && (!s->pos_vel || !s->att_fuel || !s->platform || !s->steps || !s->episode || !s->flags)) return DD_E_NULL;
    if ((reinterpret_cast<uintptr_t>(s->pos_vel) | reinterpret_cast<uintptr_t>(s->att_fuel)) & 15u) return DD_E_ALIGN;
    if (reinterpret_cast<uintptr_t>(s->platform) & (2 * sizeof(R) - 1)) return DD_E_ALIGN;
    a = KArgs<R>{};
    a.pos_vel = (R*)s->pos_vel; a.att_fuel = (R*)s->att_fuel; a.platform = (R*)s->platform;
    a.steps = s->steps; a.episode = s->episode; a.flags = s->flags;
    a.n = n; a.seed = c->seed; a.env_id_base = c->env_id_base;
    a.max_steps = c->max_steps; a.obs_stride = DD_OBS_DIM;
    a.rand_drone = c->randomize_drone; a.rand_platform = c->randomize_platform;
    a.k = make_consts<R>(*p);
    return 0;
}

static inline int check_stride(int32_t stride) {
    return (stride == DD_OBS_DIM || stride == kMaxObsStride) ? 0 : DD_E_RANGE;
}

template <typename R>
static int reset_impl(const DDState* s, const DDParams* p, const DDEnvConfig* c, const uint8_t* mask,
                      void* obs, int32_t obs_stride, int64_t n, cudaStream_t st)
{
    KArgs<R> a;
    if (int rc = fill_args(a, s, p, c, n)) return rc;
    if (obs) { if (int rc = check_stride(obs_stride)) return rc; a.obs_stride = obs_stride; }
    a.obs = (R*)obs;
    if (n == 0) return 0;
    reset_kernel<R><<<grid_for(n), kBlock, 0, st>>>(a, mask);
    return (int)cudaGetLastError();
}

template <typename R>
static int step_impl(const DDState* s, const DDParams* p, const DDEnvConfig* c, const uint8_t* actions,
                     void* obs, int32_t obs_stride, void* reward, uint8_t* done_flags, void* final_obs,
                     uint64_t* stats, int64_t n, cudaStream_t st)
{
    KArgs<R> a;
    if (int rc = fill_args(a, s, p, c, n)) return rc;
    if (!actions && n > 0) return DD_E_NULL;
    if (obs || final_obs) { if (int rc = check_stride(obs_stride)) return rc; a.obs_stride = obs_stride; }
    a.actions = actions; a.obs = (R*)obs; a.reward = (R*)reward; a.final_obs = (R*)final_obs;
    a.done_flags = done_flags; a.stats = (unsigned long long*)stats;
    if (n == 0) return 0;
    const int g = grid_for(n);
    const bool au = c->auto_reset != 0, ob = obs != nullptr;
    if (au && ob) step_kernel<R, true, true><<<g, kBlock, 0, st>>>(a);
    else if (au) step_kernel<R, true, false><<<g, kBlock, 0, st>>>(a);
    else if (ob) step_kernel<R, false, true><<<g, kBlock, 0, st>>>(a);
    else step_kernel<R, false, false><<<g, kBlock, 0, st>>>(a);
    return (int)cudaGetLastError();
}

template <typename R>
static int rollout_impl(const DDState* s, const DDParams* p, const DDEnvConfig* c, int32_t policy,
                        const uint8_t* actions_tn, uint32_t t0, int32_t T, void* reward_tn, uint8_t* done_tn,
                        void* obs_tn, int32_t obs_stride, uint64_t* stats, int64_t n, cudaStream_t st)
{
    RArgs<R> ra{};
    if (int rc = fill_args(ra.a, s, p, c, n)) return rc;
    if (policy < DD_POLICY_TRACE || policy > DD_POLICY_BANGBANG) return DD_E_RANGE;
    if (policy == DD_POLICY_TRACE && !actions_tn) return DD_E_NULL;
    if (T < 0) return DD_E_RANGE;
    if (obs_tn) { if (int rc = check_stride(obs_stride)) return rc; ra.a.obs_stride = obs_stride; }
    ra.a.stats = (unsigned long long*)stats;
    ra.actions_tn = actions_tn; ra.reward_tn = (R*)reward_tn; ra.done_tn = done_tn; ra.obs_tn = (R*)obs_tn;
    ra.t0 = t0; ra.T = T; ra.policy = policy; ra.auto_reset = c->auto_reset;
    if (n == 0 || T == 0) return 0;
    if (obs_tn) rollout_kernel<R, true><<<grid_for(n), kBlock, 0, st>>>(ra);
    else rollout_kernel<R, false><<<grid_for(n), kBlock, 0, st>>>(ra);
    return (int)cudaGetLastError();
}

}  // namespace dd

extern "C" {

int dd_abi_version(void) { return DD_ABI_VERSION; }

void dd_default_params(DDParams* p)
{
    if (!p) return;
    p->width = 800; p->height = 600;
    p->gravity = 0.3; p->drag = 0.99; p->angular_drag = 0.95;
    p->drone_height = 20;
    p->main_thrust = 0.6; p->side_thrust = 0.3;
    p->max_fuel = 1000.0; p->fuel_main = 2.0; p->fuel_side = 1.0;
    p->platform_w = 100; p->platform_h = 20;
    p->land_speed = 3.0; p->land_angle = 20.0;
    p->oob_margin = 50; p->ground_margin = 50;
    p->r_land = 100.0; p->r_crash = -100.0; p->r_fuel = -50.0; p->r_oob = -50.0; p->r_step = -0.1;
    p->shape_offset = 500; p->shape_div = 5000;
    p->start_x = 400; p->start_y = 100;
    p->plat_default_x = 400; p->plat_default_y = 500;
    p->spawn_x_min = 100; p->spawn_x_count = 601;
    p->spawn_y_min = 50; p->spawn_y_count = 201;
    p->plat_x_min = 100; p->plat_x_count = 600;
    p->plat_y_min = 100; p->plat_y_count = 450;
    p->vel_norm = 10.0; p->angle_norm = 180.0; p->angvel_norm = 10.0;
}

const char* dd_error_string(int code)
{
    switch (code) {
        case 0: return "ok";
        case DD_E_NULL: return "null pointer argument";
        case DD_E_RANGE: return "argument out of range";
        case DD_E_DTYPE: return "unknown dtype";
        case DD_E_ALIGN: return "buffer not sufficiently aligned";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}

int dd_reset(const DDState* s, const DDParams* p, const DDEnvConfig* c, const uint8_t* mask,
             void* obs, int32_t obs_stride, int64_t n, void* stream)
{
    if (!s) return DD_E_NULL;
    if (s->dtype == DD_F32) return dd::reset_impl<float>(s, p, c, mask, obs, obs_stride, n, (cudaStream_t)stream);
    if (s->dtype == DD_F64) return dd::reset_impl<double>(s, p, c, mask, obs, obs_stride, n, (cudaStream_t)stream);
    return DD_E_DTYPE;
}

int dd_step(const DDState* s, const DDParams* p, const DDEnvConfig* c, const uint8_t* actions,
            void* obs, int32_t obs_stride, void* reward, uint8_t* done_flags, void* final_obs,
            uint64_t* stats, int64_t n, void* stream)
{
    if (!s) return DD_E_NULL;
    if (s->dtype == DD_F32)
        return dd::step_impl<float>(s, p, c, actions, obs, obs_stride, reward, done_flags, final_obs, stats, n, (cudaStream_t)stream);
    if (s->dtype == DD_F64)
        return dd::step_impl<double>(s, p, c, actions, obs, obs_stride, reward, done_flags, final_obs, stats, n, (cudaStream_t)stream);
    return DD_E_DTYPE;
}

int dd_rollout(const DDState* s, const DDParams* p, const DDEnvConfig* c, int32_t policy,
               const uint8_t* actions_tn, uint32_t t0, int32_t T, void* reward_tn, uint8_t* done_tn,
               void* obs_tn, int32_t obs_stride, uint64_t* stats, int64_t n, void* stream)
{
    if (!s) return DD_E_NULL;
    if (s->dtype == DD_F32)
        return dd::rollout_impl<float>(s, p, c, policy, actions_tn, t0, T, reward_tn, done_tn, obs_tn, obs_stride, stats, n, (cudaStream_t)stream);
    if (s->dtype == DD_F64)
        return dd::rollout_impl<double>(s, p, c, policy, actions_tn, t0, T, reward_tn, done_tn, obs_tn, obs_stride, stats, n, (cudaStream_t)stream);
    return DD_E_DTYPE;
}

int dd_fill_random_actions(uint8_t* actions_tn, uint64_t seed, uint64_t env_id_base, uint32_t t0, int32_t T,
                           int64_t n, void* stream)
{
    if (!actions_tn) return DD_E_NULL;
    if (n < 0 || T < 0) return DD_E_RANGE;
    if (n == 0 || T == 0) return 0;
    dd::fill_random_actions_kernel<<<dd::grid_for(n), dd::kBlock, 0, (cudaStream_t)stream>>>(actions_tn, seed, env_id_base, t0, T, n);
    return (int)cudaGetLastError();
}

int dd_pack_actions(const uint8_t* actions3, uint8_t* packed, int64_t n, void* stream)
{
    if (!actions3 || !packed) return DD_E_NULL;
    if (n < 0) return DD_E_RANGE;
    if (n == 0) return 0;
    dd::pack_actions_kernel<<<dd::grid_for(n), dd::kBlock, 0, (cudaStream_t)stream>>>(actions3, packed, n);
    return (int)cudaGetLastError();
}

int dd_stats_collapse(const uint64_t* stats, uint64_t* out, void* stream)
{
    if (!stats || !out) return DD_E_NULL;
    dd::stats_collapse_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const unsigned long long*)stats, (unsigned long long*)out);
    return (int)cudaGetLastError();
}

}  // extern "C"
