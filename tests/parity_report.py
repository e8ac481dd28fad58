#!/usr/bin/env python
"""Parity numbers in one JSON (run on the GPU box: ``python tests/parity_report.py > gpurun_out/parity.json``).

Not a test (the pass/fail gates are tests/test_gpu_*.py): it reports the measured distances between the CUDA path
(through the C ABI) and the float64 oracle / the golden vectors recorded from the unmodified reference, so that
the numbers behind the tolerances can be read without re-deriving them.  Test infrastructure: imports oracle/."""
import importlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "golden"))
from corpus import N_CORPUS, N_TRAJ, T_CORPUS, corpus_actions, corpus_spawns   # noqa: E402
from oracle import c_oracle as co                                              # noqa: E402

dd = importlib.import_module("reinforcement-learning-101_b200")
DEV = "cuda:0"


def _t(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a), device=DEV)
    return t if dtype is None else t.to(dtype)


def margin(o):
    rad = np.radians(o.angle)
    bx, by = o.x - 10.0 * np.sin(rad), o.y + 10.0 * np.cos(rad)
    speed = np.sqrt(o.vx ** 2 + o.vy ** 2)
    m = [np.abs(speed - 3.0), np.abs(np.abs(o.angle) - 20.0), np.abs(bx - (o.px - 50)), np.abs(bx - (o.px + 50)),
         np.abs(by - (o.py - 10)), np.abs(by - (o.py + 10)), np.abs(o.y - 550.0), np.abs(o.x + 50.0),
         np.abs(o.x - 850.0), np.abs(o.y + 50.0)]
    return np.min(np.stack(m), axis=0)


def run(dtype, n, T, spawns, actions, golden=None):
    sx, sy, spx, spy = spawns
    o = co.OracleBatch(n, randomize_drone=False, randomize_platform=False)
    o.inject(sx, sy, spx, spy)
    e = dd.BatchedDroneEnv(n, device=DEV, dtype=dtype, randomize_drone=False, randomize_platform=False, auto_reset=False)
    e.inject(_t(sx, dtype), _t(sy, dtype), _t(spx, dtype), _t(spy, dtype))
    valid = np.ones(n, bool)
    flips, worst_obs, worst_rel, worst_rew, worst_gold = [], 0.0, 0.0, 0.0, 0.0
    counters_equal = True
    Ad = _t(actions)
    for t in range(T):
        oo, orr, od = o.step(actions[t])
        obs, rew, done, info = e.step(Ad[t])
        fl = info["flags"].cpu().numpy()
        bad = valid & (fl != od)
        if bad.any():
            mg = margin(o)
            flips += [{"env": int(i), "step": t, "gpu_flags": int(fl[i]), "oracle_flags": int(od[i]),
                       "f64_threshold_margin": float(mg[i])} for i in np.nonzero(bad)[0]]
            valid &= ~bad
        counters_equal &= bool(np.array_equal(e.steps.cpu().numpy()[valid], o.steps[valid]))
        obs, rew = obs.cpu().numpy().astype(np.float64), rew.cpu().numpy().astype(np.float64)
        d = np.abs(obs - oo)[valid]
        worst_obs = max(worst_obs, float(d.max()))
        worst_rel = max(worst_rel, float((d / np.maximum(np.abs(oo[valid]), 1.0)).max()))
        worst_rew = max(worst_rew, float((np.abs(rew - orr)[valid] / np.maximum(np.abs(orr[valid]), 1.0)).max()))
        if golden is not None:
            worst_gold = max(worst_gold, float(np.abs(obs[:N_TRAJ] - golden["obs"][t])[valid[:N_TRAJ]].max()))
    out = {"envs": n, "steps": T, "episodes_ended": int(((o.flags & co.DONE) > 0).sum()) if hasattr(o, "flags") else None,
           "flag_mismatches": len(flips), "flips": flips[:16], "step_counters_equal_on_compared_envs": counters_equal,
           "max_abs_err_normalised_obs": worst_obs, "max_err_obs_relative_to_max(|ref|,1)": worst_rel,
           "max_err_reward_relative_to_max(|ref|,1)": worst_rew}
    if golden is not None:
        out["max_abs_err_vs_reference_recorded_trajectories"] = worst_gold
    return out


def main():
    rep = {"device": torch.cuda.get_device_name(0), "tolerance": "north_star: flags/counters bit-exact, fp32 |a-b| <= 1e-5*max(|b|,1)"}
    gt = np.load(os.path.join(HERE, "golden", "corpus_traj.npz"))
    sp, A = corpus_spawns(), corpus_actions()
    rep["cfg2_float64_vs_oracle_and_reference"] = run(torch.float64, N_CORPUS, T_CORPUS, sp, A, gt)
    rep["cfg2_float32_vs_oracle_and_reference"] = run(torch.float32, N_CORPUS, T_CORPUS, sp, A, gt)
    # a larger draw for the fp32 flip rate: 65,536 envs, same spawn ranges / action mix as the corpus
    g = np.random.default_rng(2026)
    n = 65536
    sp2 = (g.integers(100, 701, n).astype(np.float64), g.integers(50, 251, n).astype(np.float64),
           g.integers(100, 700, n).astype(np.float64), g.integers(100, 550, n).astype(np.float64))
    p = np.where((np.arange(n) % 2 == 0)[None, :, None], 0.5, np.array([0.5, 0.1, 0.1])[None, None, :])
    bits = (g.random((250, n, 3)) < p).astype(np.uint8)
    A2 = (bits[..., 0] | (bits[..., 1] << 1) | (bits[..., 2] << 2)).astype(np.uint8)
    rep["float32_65536_envs_x_250_steps_vs_oracle"] = run(torch.float32, n, 250, sp2, A2)
    # K5: tensor-core network against the eager fp32 outputs recorded from the reference checkpoints
    d = np.load(os.path.join(HERE, "golden", "policy_v1.npz"))
    blob = dd.PolicyBlob({k: torch.from_numpy(d[k]) for k in d.files if k.startswith("network")}, device=DEV)
    probs = dd.policy_forward(blob, torch.from_numpy(d["obs"]).to(DEV)).cpu().numpy()
    flips = (probs > 0.5) != (d["probs"] > 0.5)
    rep["K5_policy_probs_vs_eager_fp32"] = {"rows": int(d["obs"].shape[0]), "max_abs_err": float(np.abs(probs - d["probs"]).max()),
                                            "threshold_flips": int(flips.sum()),
                                            "max_|logit|_at_a_flip": float(np.abs(d["logits"])[flips].max(initial=0.0))}
    c = np.load(os.path.join(HERE, "golden", "critic_v1.npz"))
    vb = dd.ValueBlob({k: torch.from_numpy(c[k]) for k in c.files if k.startswith("network")}, device=DEV)
    v = dd.value_forward(vb, torch.from_numpy(c["obs"]).to(DEV)).cpu().numpy()
    rep["K5_critic_values_vs_eager_fp32"] = {"rows": int(c["obs"].shape[0]), "value_std": float(c["values"].std()),
                                             "max_abs_err": float(np.abs(v - c["values"]).max()),
                                             "rms_err": float(np.sqrt(((v - c["values"]) ** 2).mean()))}
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
