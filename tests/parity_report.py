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
    # K5: tensor-core network against the eager fp32 outputs recorded from the reference checkpoints, in both operand
    # formats (fp16 is what dd_policy_pack picks for these checkpoints; bf16 is the fallback for networks outside its range)
    d = np.load(os.path.join(HERE, "golden", "policy_v1.npz"))
    c = np.load(os.path.join(HERE, "golden", "critic_v1.npz"))
    sd = {k: torch.from_numpy(d[k]) for k in d.files if k.startswith("network")}
    sdc = {k: torch.from_numpy(c[k]) for k in c.files if k.startswith("network")}
    for operands in ("fp16", "bf16"):
        blob = dd.PolicyBlob(sd, device=DEV, operands=operands)
        probs = dd.policy_forward(blob, torch.from_numpy(d["obs"]).to(DEV)).cpu().numpy()
        flips = (probs > 0.5) != (d["probs"] > 0.5)
        rep[f"K5_policy_probs_vs_eager_fp32_{operands}"] = {
            "rows": int(d["obs"].shape[0]), "max_abs_err": float(np.abs(probs - d["probs"]).max()),
            "threshold_flips": int(flips.sum()), "max_|logit|_at_a_flip": float(np.abs(d["logits"])[flips].max(initial=0.0))}
        vb = dd.ValueBlob(sdc, device=DEV, operands=operands)
        v = dd.value_forward(vb, torch.from_numpy(c["obs"]).to(DEV)).cpu().numpy()
        rep[f"K5_critic_values_vs_eager_fp32_{operands}"] = {
            "rows": int(c["obs"].shape[0]), "value_std": float(c["values"].std()),
            "max_abs_err": float(np.abs(v - c["values"]).max()), "rms_err": float(np.sqrt(((v - c["values"]) ** 2).mean()))}
    rep["K5_auto_operands"] = dd.PolicyBlob(sd, device=DEV).operand_dtype
    # ... and at size: every observation row of a cfg-4 rollout (65,536 envs x 250 steps) through the eager fp32 torch networks
    from torch_reference import reference_policy
    torch.backends.cuda.matmul.allow_tf32 = False
    for operands in ("fp16", "bf16"):
        blob, vb = dd.PolicyBlob(sd, device=DEV, operands=operands), dd.ValueBlob(sdc, device=DEV, operands=operands)
        env = dd.BatchedDroneEnv(65536, device=DEV, seed=2, randomize_drone=True, randomize_platform=True, max_steps=250,
                                 auto_reset=True, dtype=torch.float32)
        env.reset()
        out = dd.policy_rollout(env, blob, 250, sample=True, want="op")
        rows, probs, vals = out["obs"].view(-1, 15), out["probs"].view(-1, 3), dd.value_forward(vb, out["obs"]).view(-1)
        pol, crit = reference_policy(sd).to(DEV), reference_policy(sdc, head=1).to(DEV)
        perr = verr = flip_logit = 0.0; psq = vsq = vsc = 0.0; nflip = 0
        with torch.no_grad():
            for lo in range(0, rows.shape[0], 1 << 20):
                x = rows[lo:lo + (1 << 20)]
                pr = pol(x); dp = (probs[lo:lo + (1 << 20)] - pr).abs()
                perr = max(perr, dp.max().item()); psq += dp.double().pow(2).sum().item()
                fl = (probs[lo:lo + (1 << 20)] > 0.5) != (pr > 0.5); nflip += int(fl.sum().item())
                logit = torch.log(pr.clamp_min(1e-30)) - torch.log1p(-pr.clamp_max(1 - 1e-7))
                flip_logit = max(flip_logit, (logit.abs() * fl).max().item())
                vr = crit(x).squeeze(-1); dv = (vals[lo:lo + (1 << 20)] - vr).abs()
                verr = max(verr, dv.max().item()); vsq += dv.double().pow(2).sum().item(); vsc += vr.double().pow(2).sum().item()
        m = rows.shape[0]
        rep[f"K5_cfg4_buffer_vs_eager_fp32_{operands}"] = {
            "rows": m, "probs_max_abs_err": perr, "probs_rms_err": (psq / (3 * m)) ** 0.5, "threshold_flips": nflip,
            "flip_rate_per_action": nflip / (3 * m), "max_|logit|_at_a_flip": flip_logit,
            "critic_max_abs_err": verr, "critic_rms_err": (vsq / m) ** 0.5, "critic_value_rms": (vsc / m) ** 0.5}
        del env, out, rows, probs, vals
        torch.cuda.empty_cache()
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
