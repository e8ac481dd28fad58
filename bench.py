#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched delivery_drone simulator (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]           # ours (CUDA, sm_100a)
    python bench.py --impl reference [--gpus N] --steps K ...      # the reference's CPU path (port)
    torchrun ... bench.py --gpus N ...                             # one rank per GPU, weak scaling

Workload (BASELINE.json configs[2]): 1,048,576 drone envs per GPU, auto-reset with Philox spawn
randomisation of drone and platform, 250-step truncation, synthetic random actions (p = 0.5 per
thruster) read from a pre-generated device trace.  One "step" = one DroneGame.step for every env
of one 1M-env shard = ONE launch of the fused step kernel.  The state of one shard (~56 MB, ~116 MB
with observations) would sit in the 126 MB L2, so the GPU holds SHARDS independent 1M-env shards
and consecutive steps rotate over them: every launch streams its state from HBM (working set
SHARDS x ~116 MB > L2).  All shards are real, distinct environments (own global env ids).

Printed JSON line (rank 0): value = env-steps/s over all ranks for the gym-surface step (obs +
reward + flags written, 146 algorithmic B/env-step); `variants` carries the step-only (86 B) and
T-steps-per-launch numbers; `e2e` = the same step through HOST buffers (pinned H2D of actions, D2H
of obs/reward/flags each step); `roofline` for the dominant kernel; `cpu_baseline` = the
reference-faithful Python port stepped by os.cpu_count() processes on this box.
"""
from __future__ import annotations

import argparse
import importlib
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "env-steps/sec"
UNIT = "env-steps/s"
ENVS_PER_GPU = 1 << 20
MAX_STEPS = 250
BYTES_STEP_ONLY = 86      # read 44 + action 1, write 36 + reward 4 + flags 1
BYTES_GYM = 146           # + 60 B observation row
WORKLOAD = ("1M drone envs per GPU, auto-reset + spawn randomisation (drone+platform), max_steps 250, "
            "synthetic random actions from a device trace")


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------
# clocks sampling during the timed region (NVML in a thread; no subprocess, no GPU work)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
        0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index: int, period_s: float = 0.01):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.period, self._stop, self._thr, self.h = period_s, threading.Event(), None, None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    index = int(vis.split(",")[index])
                except ValueError:
                    pass
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:          # NVML missing: report it, do not fail the bench
            self.err = repr(e)

    def _one(self):
        nv = self.nv
        self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        for bit, name in self.REASONS.items():
            if r & bit and name != "gpu_idle":
                self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._one()
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.h is not None:
            self._stop.clear()
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": getattr(self, "err", "nvml unavailable")}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ---------------------------------------------------------------------------------------------
# the CPU reference arm: reference-faithful Python port, one process per host core
# ---------------------------------------------------------------------------------------------
def cpu_port_multiprocess(ticks: int, warmup_ticks: int, procs: int | None = None, games: int = 6):
    """BASELINE.json configs[0]: 6 DroneGame instances per process, random actions, 250-step cap.
    Returns (env_steps_per_s aggregate, per-process mean, procs, seconds)."""
    import multiprocessing as mp
    from oracle.drone_port import cpu_rollout_worker
    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        res = pool.map(cpu_rollout_worker, [(p, ticks, games, MAX_STEPS, warmup_ticks) for p in range(procs)])
    steps = sum(n for n, _ in res)
    slowest = max(s for _, s in res)
    per_proc = sum(n / s for n, s in res) / procs
    return steps / slowest, per_proc, procs, slowest


def c_oracle_threads(n: int, T: int, threads: int | None = None):
    """The C float64 restatement (oracle/drone_oracle.c) on all host cores: env-steps/s."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import c_oracle as co
    threads = threads or os.cpu_count() or 1
    b = co.OracleBatch(n, seed=0, randomize_drone=True, randomize_platform=True, max_steps=MAX_STEPS, auto_reset=True)
    b.reset()
    cuts = [n * i // threads for i in range(threads + 1)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:     # ctypes releases the GIL inside the C call
        list(ex.map(lambda k: b.rollout(T, policy=co.POL_RANDOM, lo=cuts[k], hi=cuts[k + 1]), range(threads)))
    return n * T / (time.perf_counter() - t0), threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    K, W = args.steps, args.warmup
    # one "step" = `tps` ticks of (6 games x P processes); sized so K steps take ~10 s of CPU work
    rate_guess = 7000.0                                   # ticks/s/process of the Python port
    tps = max(1, int(math.ceil(10.0 * rate_guess / max(K, 1))))
    tps = min(tps, max(1, int(120.0 * rate_guess / max(K + W, 1))))   # never more than ~2 min in total
    agg, per_proc, procs, secs = cpu_port_multiprocess(K * tps, W * tps)
    sample = (f"{procs} processes x 6 PortDroneGame (reference-faithful Python port of DroneGame), random actions, "
              f"250-step cap, {K} steps x {tps} ticks x {6 * procs} envs = {K * tps * 6 * procs} env-steps")
    line = {
        "impl": "reference", "metric": METRIC, "value": agg, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": secs / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "reference_sample": sample},
        "cpu_baseline": {"value": agg, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample,
                         "per_process": per_proc},
        "e2e": {"value": agg, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------
# ours
# ---------------------------------------------------------------------------------------------
def _ncu_traffic(n: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of step_kernel<float,AUTO,OBS> from the committed
    ncu capture (profiles/r01_traffic_steady.json; ncu cannot run inside the bench).  Only valid for the
    launch size it was captured at."""
    path = os.path.join(ROOT, "profiles", "r01_traffic_steady.json")
    try:
        with open(path) as f:
            d = json.load(f)
        k = d["kernels"]["void step_kernel<float, 1, 1, 1, 256>(KArgs<T1>)"]
        if d["algorithmic_bytes_per_launch"]["step_kernel<float,AUTO,OBS>"] != BYTES_GYM * n:
            return None, f"profiles/r01_traffic_steady.json was captured at another launch size than {n} envs"
        return k["dram_bytes_per_launch"], "profiles/r01_traffic_steady.json (ncu, steady state, --cache-control none)"
    except (OSError, KeyError, ValueError):
        return None, "no ncu capture committed"


def socket_shim_rate(dev, games: int = 6, seconds: float = 2.0):
    """Context number: the reference's TCP/JSON protocol served from the batched env -- 6 games,
    one client stepping them round-robin with random actions, reset on done, like the notebooks do.  The
    reference documents 300-500 steps/s for its own server (SOCKET_API.md:373); measured 58.5 steps/s at its
    default --fps 60 and ~6.7e3 with the fps cap lifted."""
    import random
    compat = importlib.import_module("reinforcement-learning-101_b200.compat")
    pool = compat.DroneGamePool(games, device=dev, seed=0, randomize_drone=True, randomize_platform=True)
    srv = compat.DroneSocketServer(pool, host="127.0.0.1", port=0)
    srv.start(background=True)
    cli = compat.DroneGameClient(host="127.0.0.1", port=srv.port, timeout=10.0)
    rng = random.Random(0)
    lat = []
    try:
        for g in range(games):
            cli.reset(g)
        t_end = time.perf_counter() + seconds
        n = 0
        while time.perf_counter() < t_end:
            g = n % games
            t0 = time.perf_counter()
            _, _, done, _ = cli.step({"main_thrust": rng.getrandbits(1), "left_thrust": rng.getrandbits(1),
                                      "right_thrust": rng.getrandbits(1)}, g)
            lat.append(time.perf_counter() - t0)
            if done:
                cli.reset(g)
            n += 1
    finally:
        cli.close()
        srv.stop()
    lat.sort()
    return {"steps_per_s": len(lat) / sum(lat), "median_step_latency_ms": lat[len(lat) // 2] * 1e3,
            "p99_step_latency_ms": lat[int(len(lat) * 0.99)] * 1e3, "games": games, "steps": len(lat),
            "what": "compat.DroneSocketServer + DroneGameClient over 127.0.0.1 (the reference's wire protocol), one request per step"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    dd = importlib.import_module("reinforcement-learning-101_b200")

    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout to the one JSON line
    rank, local, ws = dd.init_from_env("nccl")
    if ws != args.gpus and ws > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={ws}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpus = dd.bind_to_gpu_numa(local) if ws > 1 and not args.no_numa_bind else None   # host buffers next to the GPU
    K, W, S, n = args.steps, args.warmup, args.shards, args.envs
    peak_gbs, peak_src = _peaks()

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms: float) -> float:
        if ws == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- SHARDS independent 1M-env shards per GPU; global env ids are unique over the whole job ----
    envs = [dd.BatchedDroneEnv(n, device=dev, seed=0, randomize_drone=True, randomize_platform=True,
                               max_steps=MAX_STEPS, auto_reset=True, dtype=torch.float32,
                               env_id_base=(rank * S + s) * n, launch_flags=args.launch_flags) for s in range(S)]
    for e in envs:
        e.reset()
    TRACE = 64                                            # steps of pre-generated actions per shard
    traces = [e.random_actions(TRACE) for e in envs]
    launches = 0

    def eager_steps(k0: int, k: int, want_obs: bool, only_chain=None):
        for j in range(k0, k0 + k):
            s = j % S
            if only_chain is not None and s % C_ != only_chain:
                continue
            envs[s].step_raw(traces[s][(j // S) % TRACE], want_obs=want_obs)

    def timed(fn_warm, fn_run):
        fn_warm()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn_run()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    # CUDA graph of G consecutive steps (a whole number of shard rotations) removes the host launch cost
    G = S * max(1, min(TRACE, 96 // S))

    # Shards are independent environments, so their step launches need no mutual ordering: the graph
    # has C_ parallel chains (shard s on chain s % C_).  Each shard's own steps stay stream-ordered;
    # kernels of different chains overlap, which hides the fill / drain of one ~15 us launch behind
    # the next (a single chain leaves ~3 us of pipeline drain between dependent launches).
    C_ = max(1, min(args.chains, S))
    chain_streams = [torch.cuda.Stream(dev) for _ in range(C_)]

    def make_graph(want_obs: bool):
        g = torch.cuda.CUDAGraph()
        side = chain_streams[0]
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            eager_steps(0, S, want_obs)                   # warm the capture stream
            with torch.cuda.graph(g, stream=side):
                for c in range(1, C_):                    # fork
                    chain_streams[c].wait_stream(side)
                for c in range(C_):
                    with torch.cuda.stream(chain_streams[c]):
                        eager_steps(0, G, want_obs, only_chain=c)
                for c in range(1, C_):                    # join
                    side.wait_stream(chain_streams[c])
        torch.cuda.current_stream(dev).wait_stream(side)
        return g

    def run_steps(g, k: int, want_obs: bool):
        q, r = divmod(k, G)
        for _ in range(q):
            g.replay()
        eager_steps(0, r, want_obs)

    sampler = ClockSampler(local)
    results = {}
    for name, want_obs in (("gym_step", True), ("step_only", False)):
        g = make_graph(want_obs)
        if name == "gym_step":
            sampler.start()
        ms = timed(lambda: run_steps(g, max(W, 3), want_obs), lambda: run_steps(g, K, want_obs))
        if name == "gym_step":
            sampler.stop()
        launches += K
        results[name] = ms
        del g
    # if the timed region was too short for NVML to see it, sample clocks over an untimed re-run
    if sampler.h is not None and len(sampler.samples) < 5:
        g = make_graph(True)
        sampler.start()
        t_end = time.perf_counter() + 0.5
        while time.perf_counter() < t_end:
            run_steps(g, G * 8, True)
            torch.cuda.synchronize(dev)
        sampler.stop()
        del g

    # ---- T-steps-per-launch rollout kernel (state in registers; Philox actions in-kernel) ----
    T_ROLL = 50
    reps = max(1, K // T_ROLL // S) * S

    def run_rollouts(k):
        for j in range(k):
            envs[j % S].rollout(T_ROLL, "random", t0=(j // S) * T_ROLL)
    ms_roll = timed(lambda: run_rollouts(S), lambda: run_rollouts(reps))
    launches += reps

    def run_rollouts_bb(k):                                 # cfg 3 (ii): fixed policy main = vy > 1.5, computed in-kernel
        for j in range(k):
            envs[j % S].rollout(T_ROLL, "bangbang", t0=(j // S) * T_ROLL)
    ms_roll_bb = timed(lambda: run_rollouts_bb(S), lambda: run_rollouts_bb(reps))
    launches += reps

    # ---- K5: fused policy rollout, BASELINE configs[3]: 65,536 envs x 250 steps, policy MLP in-kernel ----
    k5 = None
    fix = os.path.join(ROOT, "tests", "golden", "policy_v1.npz")
    cfix_path = os.path.join(ROOT, "tests", "golden", "critic_v1.npz")
    if not args.no_policy and not (os.path.exists(fix) and os.path.exists(cfix_path)):
        k5 = {"skipped": "checkpoint fixtures tests/golden/policy_v1.npz / critic_v1.npz not found"}   # same on every rank
    elif not args.no_policy:
        import numpy as np
        d = np.load(fix)
        sd = {kk: torch.from_numpy(d[kk]) for kk in d.files if kk.startswith("network")}
        blob = dd.PolicyBlob(sd, device=dev)
        NP, TP = args.policy_envs, 250
        penv = dd.BatchedDroneEnv(NP, device=dev, seed=0, randomize_drone=True, randomize_platform=True,
                                  max_steps=MAX_STEPS, auto_reset=True, dtype=torch.float32, env_id_base=rank * NP)
        penv.reset()
        pbuf = dd.policy_rollout(penv, blob, TP, sample=True, want="arldo")      # allocates the rollout buffers
        reps_p = max(2, min(10, K // 1000))

        def run_policy(kk):
            for j in range(kk):
                dd.policy_rollout(penv, blob, TP, sample=True, t0=j * TP, want="arldo", out=pbuf)
        ms_pol = timed(lambda: run_policy(1), lambda: run_policy(reps_p))
        launches += reps_p
        FLOP_STEP = 2 * (15 * 128 + 128 * 128 + 128 * 64 + 64 * 3)              # 53,376
        steps_s = NP * TP * reps_p * ws / (ms_pol * 1e-3)
        k5 = {"value": steps_s, "envs_per_gpu": NP, "steps_per_launch": TP, "ms_per_launch": ms_pol / reps_p,
              "mlp_tflops": steps_s / ws * FLOP_STEP / 1e12, "flop_per_env_step": FLOP_STEP,
              "buffers": "obs[T,N,15] f32, action u8, logp f32, reward f32, done u8 (70 B/env-step)",
              "policy": "drone_policy_v1 (27,651 params), Bernoulli sampling, bf16 tcgen05 MMA / fp32 accumulate",
              "stats": penv.stats(reduce=ws > 1)}
        # 65,536 envs are 512 tiles of 128 = 128 CTAs: 20 of the 148 SMs idle.  One tile set per SM for reference.
        NF = 148 * 512
        fenv = dd.BatchedDroneEnv(NF, device=dev, seed=0, randomize_drone=True, randomize_platform=True,
                                  max_steps=MAX_STEPS, auto_reset=True, dtype=torch.float32, env_id_base=rank * NF)
        fenv.reset()
        fbuf = dd.policy_rollout(fenv, blob, TP, sample=True, want="arldo")

        def run_policy_full(kk):
            for j in range(kk):
                dd.policy_rollout(fenv, blob, TP, sample=True, t0=j * TP, want="arldo", out=fbuf)
        ms_full = timed(lambda: run_policy_full(1), lambda: run_policy_full(reps_p))
        launches += reps_p
        k5["all_148_sms"] = {"envs_per_gpu": NF, "ms_per_launch": ms_full / reps_p,
                             "value": NF * TP * reps_p * ws / (ms_full * 1e-3),
                             "mlp_tflops": NF * TP * reps_p / (ms_full * 1e-3) * FLOP_STEP / 1e12}
        del fenv, fbuf
        # what follows the rollout in the PPO loop (Actor_Critic_PPO.ipynb, PHASE 2): critic values over the stored
        # states + bootstrap row, compute_gae, advantage normalisation -- all on the rollout buffers in HBM
        cfix = np.load(cfix_path)
        vblob = dd.ValueBlob({kk: torch.from_numpy(cfix[kk]) for kk in cfix.files if kk.startswith("network")}, device=dev)
        final_obs = penv.observe().clone()
        vals = dd.rollout_values(vblob, pbuf["obs"], final_obs)
        dones = (pbuf["done"] != 0).to(torch.uint8)
        adv = dd.gae(pbuf["reward"], vals, dones)
        nadv = torch.empty_like(adv)

        def run_ppo_tail(kk):
            for _ in range(kk):
                dd.rollout_values(vblob, pbuf["obs"], final_obs)
                dd.gae(pbuf["reward"], vals, dones, out=adv)
                dd.normalize_advantages(adv, reduce=ws > 1, out=nadv)
        ms_tail = timed(lambda: run_ppo_tail(1), lambda: run_ppo_tail(reps_p))
        launches += 5 * reps_p
        # the same through HOST-side inputs / outputs, as one PPO iteration sees it: policy + critic weights come from
        # host state_dicts (packed and uploaded every iteration, like after an optimiser step), the rollout buffers
        # stay in HBM for the learner, the episode statistics and advantage moments go back to the host
        sd_host = {kk: v.clone() for kk, v in sd.items()}
        sdc_host = {kk: torch.from_numpy(cfix[kk]) for kk in cfix.files if kk.startswith("network")}
        t_it = []
        for it in range(3 + reps_p):
            barrier()
            t0 = time.perf_counter()
            b_it = dd.PolicyBlob(sd_host, device=dev)
            vb_it = dd.ValueBlob(sdc_host, device=dev)
            penv.reset_stats()
            dd.policy_rollout(penv, b_it, TP, sample=True, t0=(100 + it) * TP, want="arldo", out=pbuf)
            dd.rollout_values(vb_it, pbuf["obs"], penv.observe())
            dd.gae(pbuf["reward"], vals, dones, out=adv)
            dd.normalize_advantages(adv, reduce=ws > 1, out=nadv)
            st_it = penv.stats(reduce=ws > 1)                  # device -> host: synchronises
            t_it.append(time.perf_counter() - t0)
        launches += 9 * len(t_it)
        it_s = max_over_ranks(sum(t_it[3:]) / reps_p * 1e3) * 1e-3
        k5["ppo_iteration_e2e"] = {"what": "host state_dicts -> pack + upload (policy, critic) -> fused rollout -> critic values -> GAE -> "
                                           "advantage normalisation -> episode statistics back on the host; wall clock per iteration",
                                   "ms": it_s * 1e3, "env_steps_per_s": NP * TP * ws / it_s,
                                   "landing_rate": st_it["landing_rate"]}
        k5["ppo_data_path"] = {"what": "critic values [T+1,N] (tcgen05, persistent forward) + GAE + advantage normalisation on the rollout buffers",
                               "ms": ms_tail / reps_p, "samples_per_s": NP * TP * reps_p * ws / (ms_tail * 1e-3),
                               "rollout_plus_tail_ms": (ms_pol + ms_tail) / reps_p}

    # ---- BASELINE configs[4]: curriculum sweep 75 -> 250, 2M envs per GPU, stats all-reduced over ranks ----
    cur = None
    if not args.no_curriculum:
        NC = args.curriculum_envs
        cenv = dd.BatchedDroneEnv(NC, device=dev, seed=0, randomize_drone=True, randomize_platform=True,
                                  auto_reset=False, dtype=torch.float32, env_id_base=rank * NC)
        caps = dd.step_schedule(8, 75, 250).tolist()
        dd.collect_episodes(cenv, caps[0], policy="bangbang", reduce=ws > 1)           # warm-up
        stages = []

        def run_sweep():
            stages[:] = dd.curriculum_sweep(cenv, caps, policy="bangbang", reduce=ws > 1)
        ms_cur = timed(lambda: None, run_sweep)
        launches += 2 * len(caps)
        cur = {"envs_total": NC * ws, "caps": caps, "ms_total": ms_cur,
               "env_steps_per_s": sum(s_["env_steps"] for s_ in stages) / (ms_cur * 1e-3),
               "policy": "bang-bang (main = vy > 1.5), one episode per env per stage, freeze after done",
               "stages": [{k_: s_[k_] for k_ in ("max_steps", "success_rate", "avg_reward", "avg_steps")} for s_ in stages]}
        del cenv

    # ---- e2e: HOST buffers in, HOST buffers out, every step ----
    io = [envs[s].make_host_io() for s in range(min(S, 2))]
    host_trace = traces[0][:TRACE].cpu().pin_memory()
    Ke = max(3, min(K, args.e2e_steps))

    def run_e2e(k):
        for j in range(k):
            b = io[j % len(io)]
            b["actions"].copy_(host_trace[j % TRACE])     # host-side: the caller's actions of this step
            envs[j % S].step_host(b, chunks=args.e2e_chunks)
    t0 = None

    def run_e2e_timed():
        run_e2e(Ke)
    ms_e2e = timed(lambda: run_e2e(3), run_e2e_timed)
    launches += Ke * max(1, args.e2e_chunks)
    h2d = n * 1
    d2h = n * (envs[0].obs_stride * 4 + 4 + 1)

    stats = None
    for e in envs[:1]:
        stats = e.stats(reduce=ws > 1)

    if rank == 0:
        ms = results["gym_step"]
        value = n * K * ws / (ms * 1e-3)
        per_launch_s = ms * 1e-3 / K
        achieved = n * BYTES_GYM / per_launch_s / 1e9
        traffic, traffic_src = _ncu_traffic(n)
        so_ms = results["step_only"]
        so_achieved = n * BYTES_STEP_ONLY / (so_ms * 1e-3 / K) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ws, "steps": K, "warmup": max(W, 3),
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu_per_step": n, "shards_per_gpu": S,
                       "l2": f"inputs larger than L2: steps rotate over {S} independent {n}-env shards "
                             f"({S} x ~{n * 116 >> 20} MiB state+outputs > 126 MB L2)",
                       "launch": f"CUDA graph of {G} step launches, replayed, {C_} parallel chain(s) over independent shards; launch_flags={args.launch_flags:#x}", "parallelism": f"env-sharded x{ws}, no per-step comms"},
            "roofline": {"bound": "hbm", "kernel": "step_kernel<float,AUTO,OBS>", "achieved": achieved, "peak": peak_gbs,
                         "unit": "GB/s", "frac": achieved / peak_gbs, "frac_of_nominal_8000_gbs": achieved / 8000.0,
                         "traffic": traffic, "traffic_unit": "bytes per launch",
                         "traffic_source": traffic_src, "algorithmic_bytes_per_launch": BYTES_GYM * n, "peak_source": peak_src,
                         "algorithmic_bytes_per_env_step": BYTES_GYM, "env_steps_per_launch": n},
            "variants": {
                "step_only": {"value": n * K * ws / (so_ms * 1e-3), "ms_per_step": so_ms / K,
                              "algorithmic_bytes_per_env_step": BYTES_STEP_ONLY, "achieved_gbs": so_achieved,
                              "frac": so_achieved / peak_gbs},
                "rollout_T50_fixed_policy_bangbang": {"value": n * T_ROLL * reps * ws / (ms_roll_bb * 1e-3),
                                                      "ms_per_launch": ms_roll_bb / reps, "steps_per_launch": T_ROLL},
                "rollout_T50_in_kernel_actions": {"value": n * T_ROLL * reps * ws / (ms_roll * 1e-3),
                                                  "ms_per_launch": ms_roll / reps, "steps_per_launch": T_ROLL},
                "fused_policy_rollout": k5,
                "curriculum_sweep": cur,
            },
            "e2e": {"value": n * Ke * ws / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": Ke, "ms_per_step": ms_e2e / Ke,
                    "api": f"BatchedDroneEnv.step_host(chunks={args.e2e_chunks}) (pinned host actions in; obs, reward, flags out)",
                    "pcie_gbs": (h2d + d2h) / (ms_e2e / Ke * 1e-3) / 1e9,
                    "bound": "PCIe: 65 B per env-step device->host (obs 60 + reward 4 + flags 1)",
                    "rank0_cpu_affinity": None if cpus is None else f"{len(cpus)} CPUs local to the GPU (NVML)"},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "episode_stats_shard0": stats,
        }
        if ws == 1 and not args.no_socket:
            try:
                line["socket_shim"] = socket_shim_rate(dev)
            except Exception as exc:                       # context only: never fail the bench line over it
                line["socket_shim"] = {"error": repr(exc)}
        if ws == 1 and not args.no_cpu_baseline:
            ticks = args.cpu_ticks
            agg, per_proc, procs, secs = cpu_port_multiprocess(ticks, 200)
            c_rate, c_thr = c_oracle_threads(1 << 16, 250)
            line["cpu_baseline"] = {
                "value": agg, "unit": UNIT, "cores": procs, "kind": "port",
                "sample": f"{procs} processes x 6 PortDroneGame x {ticks} ticks (random actions, 250-step cap), {secs:.1f} s",
                "per_process": per_proc,
                "c_oracle": {"value": c_rate, "threads": c_thr, "sample": "65,536 envs x 250 steps, float64 C restatement"},
            }
        print(json.dumps(line), flush=True)
    if ws > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=12000)
    ap.add_argument("--warmup", type=int, default=600)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="envs per shard (= per step launch)")
    ap.add_argument("--shards", type=int, default=6, help="independent shards per GPU that steps rotate over")
    ap.add_argument("--chains", type=int, default=2, help="parallel launch chains in the CUDA graph (shard s -> chain s %% chains)")
    ap.add_argument("--e2e-steps", type=int, default=300)
    ap.add_argument("--e2e-chunks", type=int, default=1, help="step_host pipelines the step over this many env slices (D2H of slice k overlaps H2D + kernel of slice k+1)")
    ap.add_argument("--cpu-ticks", type=int, default=250000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin each rank to its GPU's local CPUs (N > 1)")
    ap.add_argument("--no-socket", action="store_true", help="skip the socket-shim context measurement")
    ap.add_argument("--no-policy", action="store_true", help="skip the fused policy rollout variant (K5)")
    ap.add_argument("--no-curriculum", action="store_true", help="skip the curriculum sweep variant (cfg 5)")
    ap.add_argument("--curriculum-envs", type=int, default=1 << 21, help="envs per GPU in the curriculum sweep")
    ap.add_argument("--policy-envs", type=int, default=65536, help="envs per GPU for the K5 variant (BASELINE configs[3])")
    ap.add_argument("--launch-flags", type=lambda v: int(v, 0), default=0x01,
                    help="DD_LAUNCH_* bits (include/drone_b200.h): 0x01 PDL, 0x10 CTA 128, 0x20 CTA 512")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
