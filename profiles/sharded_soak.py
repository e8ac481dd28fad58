#!/usr/bin/env python
"""Soak test of the product launch path: ShardedDroneEnv.run with random k for a few thousand calls, graph cache small
enough to evict constantly, max_steps changes (graphs dropped) and step_all interleaved; the graph path must stay
bit-identical to the eager path the whole way.   python profiles/sharded_soak.py [--calls 3000]"""
import argparse
import importlib
import os
import random
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dd = importlib.import_module("reinforcement-learning-101_b200")
ap = argparse.ArgumentParser(); ap.add_argument("--calls", type=int, default=3000); a = ap.parse_args()
dev = "cuda:0"
kw = dict(seed=3, randomize_drone=True, randomize_platform=True, max_steps=60, auto_reset=True, dtype=torch.float32)
S, N, L = 5, 4099, 7
g = dd.ShardedDroneEnv(S, N, device=dev, chains=2, trace_len=L, max_graphs=12, **kw)
e = dd.ShardedDroneEnv(S, N, device=dev, chains=3, trace_len=L, use_graphs=False, **kw)
g.reset(); e.reset()
e.set_trace(g.random_trace().clone())
rng = random.Random(0)
t0 = time.time()
checks = 0
for c in range(a.calls):
    r = rng.random()
    if r < 0.9:
        k = rng.choice([1, 2, 3, 5, 7, 20, 35, 36, 70, 71])
        g.run(k, want_obs=(c % 3 != 0)); e.run(k, want_obs=(c % 3 != 0))
    elif r < 0.97:
        act = torch.randint(0, 8, (S, N), dtype=torch.uint8, device=dev)
        g.step_all(act); e.step_all(act)
    else:
        ms = rng.choice([40, 60, 90])
        g.max_steps = ms; e.max_steps = ms
    if c % 250 == 249:
        g.join(); e.join(); torch.cuda.synchronize()
        for s in range(S):
            sa, sb = g.shards[s].get_state(), e.shards[s].get_state()
            for key in sa:
                if key != "prev_dist":
                    assert torch.equal(sa[key], sb[key]), (c, s, key)
        assert g.stats() == e.stats(), c
        checks += 1
print(f"soak ok: {a.calls} calls, {g.t} launches, {checks} full comparisons, {g.graph_replays} graph replays, "
      f"{g.graphs_cached} pieces cached (cap 12), {time.time() - t0:.1f} s; stats {g.stats()}")
