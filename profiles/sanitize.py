#!/usr/bin/env python
"""One small launch of every kernel family, for compute-sanitizer:
    compute-sanitizer --tool memcheck  python profiles/sanitize.py
    compute-sanitizer --tool racecheck python profiles/sanitize.py
K1 step (fp32 + fp64, obs via the TMA bulk store and via the ragged fallback), K2 reset, the T-step rollout with
shaped rewards, K5 fused policy rollout (tcgen05 / TMEM / mbarrier / TMA store) in both operand formats, the forward-only
critic (TMA loads), GAE (cp.async ring and register kernels) + moments + normalise + discounted returns, stats collapse."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dd = importlib.import_module("reinforcement-learning-101_b200")
dev = "cuda:0"
kw = dict(seed=3, randomize_drone=True, randomize_platform=True, max_steps=20, auto_reset=True)
for dtype, n in ((torch.float32, 2048), (torch.float32, 1000), (torch.float64, 515)):
    e = dd.BatchedDroneEnv(n, device=dev, dtype=dtype, want_final_obs=True, **kw)
    e.reset()
    a = e.random_actions(6)
    for t in range(6):
        e.step_raw(a[t])
    e.step_raw(a[0], want_obs=False)
    obs = torch.empty(8, n, 15, dtype=dtype, device=dev); shp = torch.empty(8, n, dtype=dtype, device=dev)
    rew = torch.empty(8, n, dtype=dtype, device=dev); don = torch.empty(8, n, dtype=torch.uint8, device=dev)
    e.rollout(8, "random", obs_out=obs, shaped_out=shp, reward_out=rew, done_out=don)
    e.rollout(8, "bangbang")
    e.stats()
d = np.load(os.path.join(ROOT, "tests", "golden", "policy_v1.npz"))
sd = {k: torch.from_numpy(d[k]) for k in d.files if k.startswith("network")}
c = np.load(os.path.join(ROOT, "tests", "golden", "critic_v1.npz"))
sdc = {k: torch.from_numpy(c[k]) for k in c.files if k.startswith("network")}
for operands in ("fp16", "bf16"):
    blob = dd.PolicyBlob(sd, device=dev, operands=operands)
    for n in (1024, 700):                                   # TMA observation store / ragged fallback
        e = dd.BatchedDroneEnv(n, device=dev, dtype=torch.float32, **kw)
        e.reset()
        out = dd.policy_rollout(e, blob, 6, sample=True, want="arldo")
        dd.policy_rollout(e, blob, 4, sample=False, want="arldops")
    vblob = dd.ValueBlob(sdc, device=dev, operands=operands)
    dd.policy_forward(blob, out["obs"].view(-1, 15)[:1500])
    vals = dd.rollout_values(vblob, out["obs"], e.observe().clone())
T, n = 23, 640
g = torch.Generator(device=dev).manual_seed(0)
for n in (640, 641):                                        # ring kernel / register kernel
    r = torch.randn(T, n, device=dev, generator=g); v = torch.randn(T + 1, n, device=dev, generator=g)
    dn = (torch.rand(T, n, device=dev, generator=g) < 0.05).to(torch.uint8)
    m = torch.zeros(3, dtype=torch.float64, device=dev)
    adv, ret = dd.gae(r, v, dn, want_returns=True, moments=m)
    dd.normalize_advantages(adv, reduce=False, moments=m)
    dd.normalize_advantages(adv, reduce=False)
    dd.discounted_returns(r, dn)
torch.cuda.synchronize()
print("sanitize run ok")
