#!/usr/bin/env python
"""A/B timing of the rollout-tail kernels (dd_gae, dd_moments, dd_normalize, dd_discounted_returns) on BASELINE
configs[3]-sized buffers (65,536 envs x 250 steps), CUDA events, against their algorithmic HBM bytes.
    DRONE_B200_LIB=<variant .so> python profiles/gae_bench.py [--label X]
Working sets: GAE reads 147 MB + writes 66 MB per launch (> the 126 MB L2); moments / normalise rotate over four 66 MB
buffers so that no launch finds its input in L2."""
import argparse
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dd = importlib.import_module("reinforcement-learning-101_b200")
ap = argparse.ArgumentParser(); ap.add_argument("--label", default=os.environ.get("DRONE_B200_LIB", "default")); ap.add_argument("--reps", type=int, default=200); ap.add_argument("--envs", type=int, default=65536); ap.add_argument("--real", action="store_true", help="time the scan on the buffers of a real fused rollout (policy + critic checkpoints) instead of random data")
a = ap.parse_args()
dev = torch.device("cuda:0")
n, T = a.envs, 250
g = torch.Generator(device=dev).manual_seed(0)
rew = torch.randn(T, n, device=dev, generator=g); val = torch.randn(T + 1, n, device=dev, generator=g)
don = (torch.rand(T, n, device=dev, generator=g) < 0.01).to(torch.uint8)
if a.real:
    import numpy as np
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = np.load(os.path.join(root, "tests", "golden", "policy_v1.npz")); c = np.load(os.path.join(root, "tests", "golden", "critic_v1.npz"))
    blob = dd.PolicyBlob({k: torch.from_numpy(d[k]) for k in d.files if k.startswith("network")}, device=dev)
    vblob = dd.ValueBlob({k: torch.from_numpy(c[k]) for k in c.files if k.startswith("network")}, device=dev)
    env = dd.BatchedDroneEnv(n, device=dev, seed=0, randomize_drone=True, randomize_platform=True, max_steps=250, auto_reset=True)
    env.reset()
    pb = dd.policy_rollout(env, blob, T, want="arldo")
    rew = pb["reward"]; val = dd.rollout_values(vblob, pb["obs"], env.observe().clone()); don = (pb["done"] != 0).to(torch.uint8)
    print("real data: done rate %.4f, |value| mean %.1f, reward mean %.3f" % (don.float().mean().item(), val.abs().mean().item(), rew.mean().item()), file=sys.stderr)
bufs = [torch.randn(T, n, device=dev, generator=g) for _ in range(4)]
adv = torch.empty(T, n, device=dev); out2 = [torch.empty(T, n, device=dev) for _ in range(2)]
mom = torch.zeros(3, dtype=torch.float64, device=dev)
lib = dd.native.lib()
st = torch.cuda.current_stream().cuda_stream
peak = 6543.4
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, reps):
    fn(5); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(reps); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def f_gae(k):
    for _ in range(k):
        dd.gae(rew, val, don, out=adv)


def f_gae_ret(k):
    for _ in range(k):
        dd.gae(rew, val, don, out=adv, out_returns=out2[0])


mom_f = torch.zeros(3, dtype=torch.float64, device=dev)


def f_gae_mom(k):                                           # the scan + the advantage moments in the same pass
    for _ in range(k):
        dd.gae(rew, val, don, out=adv, moments=mom_f)


def f_mom(k):
    for j in range(k):
        dd.advantage_moments(bufs[j % 4], out=mom)


def f_nrm(k):
    for j in range(k):
        lib.dd_normalize(bufs[j % 4].data_ptr(), out2[j % 2].data_ptr(), mom.data_ptr(), 1e-8, n * T, st)


def f_ret(k):
    for j in range(k):
        dd.discounted_returns(bufs[j % 4], don, out=out2[j % 2])


ne = n * T
res = {"label": a.label, "envs": n, "elements": ne, "peak_gbs": peak}
for name, fn, b in (("gae", f_gae, 13.0 + 4.0 / T), ("gae_fused_moments", f_gae_mom, 13.0 + 4.0 / T), ("gae_with_returns", f_gae_ret, 17.0 + 4.0 / T), ("moments", f_mom, 4.0), ("normalize", f_nrm, 8.0),
                    ("discounted_returns", f_ret, 9.0)):
    ms = timed(fn, a.reps)
    res[name] = {"ms": ms, "gbs": ne * b / (ms * 1e-3) / 1e9, "frac": ne * b / (ms * 1e-3) / 1e9 / peak, "bytes_per_element": b}
# bit-exactness of the GAE scan against the eager torch loop of the notebook (Actor_Critic_PPO.ipynb c15:L49-53) on a slice
r_, v_, d_ = rew[:, :4096], val[:, :4096], don[:, :4096]
A = dd.gae(r_.contiguous(), v_.contiguous(), d_.contiguous())
ref = torch.zeros_like(A); gae_ = torch.zeros(4096, device=dev)
for t in reversed(range(T)):
    mask = 1.0 - d_[t].float()
    delta = r_[t] + 0.99 * v_[t + 1] * mask - v_[t]
    gae_ = delta + 0.99 * 0.95 * mask * gae_
    ref[t] = gae_
res["gae_bit_exact_vs_eager_torch_loop"] = bool(torch.equal(A, ref))
m1 = torch.zeros(3, dtype=torch.float64, device=dev); m2 = torch.zeros(3, dtype=torch.float64, device=dev)
dd.gae(rew, val, don, out=adv, moments=m1); dd.advantage_moments(adv, out=m2)
res["fused_moments_rel_err_vs_dd_moments"] = float(((m1 - m2).abs() / m2.abs().clamp_min(1e-30)).max().item())
print(json.dumps(res), flush=True)
