#!/usr/bin/env python
"""Phase timeline of the fused policy kernel (profiling build -DDD_K5_TRACE=1): clock64 stamps of CTA 0's four tiles at
12 points of 24 consecutive steps (t = 100..123), printed as cycle offsets.
    DRONE_B200_LIB=build_variants/libdd_k5_trace.so python profiles/k5_trace.py"""
import ctypes as C
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dd = importlib.import_module("reinforcement-learning-101_b200")
d = np.load(os.path.join(ROOT, "tests", "golden", "policy_v1.npz"))
blob = dd.PolicyBlob({k: torch.from_numpy(d[k]) for k in d.files if k.startswith("network")}, device="cuda:0")
env = dd.BatchedDroneEnv(65536, device="cuda:0", seed=0, randomize_drone=True, randomize_platform=True, max_steps=250, auto_reset=True)
env.reset()
buf = dd.policy_rollout(env, blob, 250, want="arldo")
dd.policy_rollout(env, blob, 250, want="arldo", out=buf)
torch.cuda.synchronize()
out = np.zeros((4, 24, 12), np.int64)
rc = dd.native.lib().dd_k5_trace_read(out.ctypes.data_as(C.c_void_p))
assert rc == 0, rc
names = ["start", "A0 written", "past bar1", "shadow1 done", "MMA1 done", "E1 done", "shadow2 done", "MMA2 done", "E2 done", "shadow3 done", "MMA3 done", "E3+head done"]
t0 = out[:, 0, 0].min()
res = {"names": names, "per_tile_mean_phase_cycles": {}, "step_period_cycles": {}}
for g in range(4):
    rel = out[g] - t0
    dur = np.diff(np.concatenate([out[g], np.roll(out[g][:, :1], -1, axis=0)], axis=1), axis=1)[:-1]     # phase durations incl. tail -> next start
    res["per_tile_mean_phase_cycles"][f"tile{g}"] = {f"{names[i]} -> {names[(i + 1) % 12]}": float(dur[:, i].mean()) for i in range(12)}
    res["step_period_cycles"][f"tile{g}"] = float(np.diff(out[g][:, 0]).mean())
    print(f"tile {g}: period {res['step_period_cycles'][f'tile{g}']:.0f} cycles; start offsets of steps 0..3 vs tile-0 start:", rel[:4, 0].tolist())
    for i in range(12):
        print(f"    {names[i]:>14} -> {names[(i + 1) % 12]:<14} {dur[:, i].mean():8.0f}  (min {dur[:, i].min():6d} max {dur[:, i].max():6d})")
# interleaving of the four tiles inside one step: all 48 stamps of step 2 sorted
ev = sorted((int(out[g, 2, i] - out[0, 2, 0]), g, names[i]) for g in range(4) for i in range(12))
print("step 102, all tiles, cycles from tile 0's step start:")
for c, g, n in ev:
    print(f"   {c:7d}  tile {g}  {n}")
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "k5_trace.json"), "w"), indent=1)
