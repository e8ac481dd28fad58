#!/usr/bin/env python
"""Stand-alone timing of K5 (the fused policy rollout, BASELINE configs[3]: 65,536 envs x 250 steps) --
the same call bench.py's `fused_policy_rollout` variant makes, without the rest of the bench, so that an
`ncu --set full -k regex:policy_rollout -c 1` capture of it takes seconds.
    python profiles/k5_bench.py [--envs 65536] [--T 250] [--reps 10] [--want arldo] [--threshold]
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dd = importlib.import_module("reinforcement-learning-101_b200")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=int(os.environ.get("K5_ENVS", 65536)))
    ap.add_argument("--T", type=int, default=250)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--want", default="arldo")
    ap.add_argument("--threshold", action="store_true")
    ap.add_argument("--operands", default="auto", choices=["auto", "bf16", "fp16"], help="16-bit format of the tensor-core operands")
    ap.add_argument("--checksum", action="store_true", help="add sums of every output buffer and of the final env state")
    ap.add_argument("--values", action="store_true", help="also time critic values over the buffer, GAE, normalisation")
    args = ap.parse_args()
    dev = "cuda:0"
    d = np.load(os.path.join(ROOT, "tests", "golden", "policy_v1.npz"))
    sd = {k: torch.from_numpy(d[k]) for k in d.files if k.startswith("network")}
    blob = dd.PolicyBlob(sd, device=dev, operands=args.operands)
    env = dd.BatchedDroneEnv(args.envs, device=dev, seed=0, randomize_drone=True, randomize_platform=True,
                             max_steps=250, auto_reset=True, dtype=torch.float32)
    env.reset()
    buf = dd.policy_rollout(env, blob, args.T, sample=not args.threshold, want=args.want)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for j in range(args.reps):
        dd.policy_rollout(env, blob, args.T, sample=not args.threshold, want=args.want, out=buf)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    steps = args.envs * args.T
    out = {"kernel": "policy_rollout_kernel", "operands": blob.operand_dtype, "lib": os.environ.get("DRONE_B200_LIB", "default"), "envs": args.envs, "T": args.T, "want": args.want,
           "ms_per_launch": ms, "env_steps_per_s": steps / ms * 1e3,
           "mlp_tflops": steps * 53376 / ms * 1e3 / 1e12, "stats": env.stats()}
    if args.checksum:                                        # A/B runs of kernel variants: identical numbers <=> identical buffers
        out["checksum"] = {k_: (float(v.double().sum()) if v.is_floating_point() else int(v.long().sum())) for k_, v in sorted(buf.items())}
        out["checksum"]["state"] = float(sum(v.double().nan_to_num().sum() for v in env.get_state().values()))
    if args.values and "obs" in buf:
        # the critic over the whole rollout buffer + bootstrap row, then GAE and advantage normalisation
        c = np.load(os.path.join(ROOT, "tests", "golden", "critic_v1.npz"))
        vblob = dd.ValueBlob({k: torch.from_numpy(c[k]) for k in c.files if k.startswith("network")}, device=dev, operands=args.operands)
        final_obs = env.observe().clone()
        vals = dd.rollout_values(vblob, buf["obs"], final_obs)
        done = (buf["done"] != 0).to(torch.uint8)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        for _ in range(args.reps):
            vals = dd.rollout_values(vblob, buf["obs"], final_obs)
        ev[1].record()
        adv = dd.gae(buf["reward"], vals, done)
        ev[1].record()
        for _ in range(args.reps):
            dd.gae(buf["reward"], vals, done, out=adv)
        ev[2].record()
        for _ in range(args.reps):
            nadv = dd.normalize_advantages(adv, reduce=False)
        ev[3].record()
        mom = torch.zeros(3, dtype=torch.float64, device=dev)
        ev2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev2[0].record()
        for _ in range(args.reps):
            dd.advantage_moments(adv, out=mom)
        ev2[1].record()
        for _ in range(args.reps):
            dd.normalize_advantages(adv, reduce=False, out=nadv)
        ev2[2].record()
        torch.cuda.synchronize()
        out["moments_only_ms"] = ev2[0].elapsed_time(ev2[1]) / args.reps
        out["normalize_into_out_ms"] = ev2[1].elapsed_time(ev2[2]) / args.reps
        out["values_ms"] = ev[0].elapsed_time(ev[1]) / args.reps
        out["values_rows_per_s"] = (steps + args.envs) / out["values_ms"] * 1e3
        out["values_mlp_tflops"] = (steps + args.envs) * 2 * (15 * 128 + 128 * 128 + 128 * 64 + 64) / out["values_ms"] * 1e3 / 1e12
        out["gae_ms"] = ev[1].elapsed_time(ev[2]) / args.reps
        out["normalize_ms"] = ev[2].elapsed_time(ev[3]) / args.reps
    print(json.dumps(out))


if __name__ == "__main__":
    main()
