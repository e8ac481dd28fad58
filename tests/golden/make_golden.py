#!/usr/bin/env python
"""Generate the golden vectors by running the UNMODIFIED reference DroneGame.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/kat.json, corpus_summary.npz, corpus_traj.npz.
The reference imports pygame at module top but never touches it headless
(game_engine.py:27), so an empty stub module is enough.
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from corpus import N_CORPUS, N_TRAJ, T_CORPUS, corpus_actions, corpus_spawns  # noqa: E402

REF = os.environ.get("DD_REFERENCE", "/root/reference")
sys.modules.setdefault("pygame", types.ModuleType("pygame"))
sys.path.insert(0, REF)
from delivery_drone.game.game_engine import DroneGame  # noqa: E402

OBS_KEYS = ("drone_x", "drone_y", "drone_vx", "drone_vy", "drone_angle", "drone_angular_vel",
            "drone_fuel", "platform_x", "platform_y", "distance_to_platform", "dx_to_platform",
            "dy_to_platform", "speed", "landed", "crashed")


def act(bits):
    return {"main_thrust": int(bits & 1), "left_thrust": int((bits >> 1) & 1), "right_thrust": int((bits >> 2) & 1)}


def summarise(g, last_r, total):
    d = g.drone
    return {"steps": int(g.steps), "last_reward": float(last_r), "total_reward": float(total),
            "x": float(d.x), "y": float(d.y), "vx": float(d.vx), "vy": float(d.vy),
            "angle": float(d.angle), "angvel": float(d.angular_velocity), "fuel": float(d.fuel),
            "landed": bool(d.landed), "crashed": bool(d.crashed)}


def run_kat(policy, cap=2000):
    g = DroneGame(render_mode=None, randomize_drone=False, randomize_platform=False)
    g.reset()
    total, r, head = 0.0, 0.0, []
    state = None
    while not g.done and g.steps < cap:
        state, r, _, info = g.step(policy(g))
        total += r
        if g.steps <= 3:
            head.append({"y": float(g.drone.y), "r": float(r)})
    out = summarise(g, r, total)
    out["head"] = head
    out["final_state"] = {k: (bool(v) if isinstance(v, (bool, np.bool_)) else float(v)) for k, v in state.items()}
    out["final_info"] = {k: float(v) for k, v in info.items()}
    # freeze-after-done (game_engine.py:107-111, BUGFIX.md:40-52)
    s2, r2, d2, i2 = g.step({"main_thrust": 1})
    out["after_done"] = {"reward": float(r2), "done": bool(d2), "steps": int(s2["steps"]),
                         "needs_reset": bool(i2.get("needs_reset", False)), "info_keys": sorted(i2.keys())}
    return out


def main():
    kat = {
        "KAT1_no_thrust": run_kat(lambda g: {}),
        "KAT2_main": run_kat(lambda g: {"main_thrust": 1}),
        "KAT3_all": run_kat(lambda g: {"main_thrust": 1, "left_thrust": 1, "right_thrust": 1}),
        "KAT4_right": run_kat(lambda g: {"right_thrust": 1}),
        "KAT5_main_right": run_kat(lambda g: {"main_thrust": 1, "right_thrust": 1}),
        "KAT6_bangbang": run_kat(lambda g: {"main_thrust": int(g.drone.vy > 1.5)}),
    }
    np.random.seed(42)
    g = DroneGame(render_mode=None, randomize_drone=True, randomize_platform=True)
    g.reset()
    kat["KAT7_seed42_spawn"] = [int(g.drone.x), int(g.drone.y), int(g.platform.x), int(g.platform.y)]
    kat["numpy"] = np.__version__
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1, sort_keys=True)

    # randomised corpus
    sx, sy, spx, spy = corpus_spawns()
    A = corpus_actions()
    N, T = N_CORPUS, T_CORPUS
    obs_sum = np.zeros((T, 15)); rew_sum = np.zeros(T); done_cnt = np.zeros(T, np.int32)
    done_step = np.zeros(N, np.int16); flags = np.zeros(N, np.uint8); total = np.zeros(N)
    final = np.zeros((N, 7))
    traj_obs = np.zeros((T, N_TRAJ, 15)); traj_rew = np.zeros((T, N_TRAJ)); traj_done = np.zeros((T, N_TRAJ), np.uint8)
    for i in range(N):
        g = DroneGame(render_mode=None, randomize_drone=False, randomize_platform=False)
        g.reset(); g.drone.reset(int(sx[i]), int(sy[i])); g.platform.reset(int(spx[i]), int(spy[i]))
        for t in range(T):
            s, r, d, info = g.step(act(A[t, i]))
            o = np.array([float(s[k]) for k in OBS_KEYS])
            obs_sum[t] += o; rew_sum[t] += r; done_cnt[t] += int(d)
            if i < N_TRAJ:
                traj_obs[t, i] = o; traj_rew[t, i] = r; traj_done[t, i] = d
            if d and done_step[i] == 0:
                done_step[i] = g.steps
        d_ = g.drone
        flags[i] = (1 if g.done else 0) | (2 if d_.landed else 0) | (4 if d_.crashed else 0)
        total[i] = g.total_reward
        final[i] = [d_.x, d_.y, d_.vx, d_.vy, d_.angle, d_.angular_velocity, d_.fuel]
    np.savez_compressed(os.path.join(HERE, "corpus_summary.npz"), obs_sum=obs_sum, rew_sum=rew_sum,
                        done_cnt=done_cnt, done_step=done_step, flags=flags, total=total, final=final)
    np.savez_compressed(os.path.join(HERE, "corpus_traj.npz"), obs=traj_obs, reward=traj_rew, done=traj_done)
    print("episodes done:", int((done_step > 0).sum()), "landed:", int(((flags & 2) > 0).sum()),
          "mean len:", float(np.where(done_step > 0, done_step, T).mean()))


if __name__ == "__main__":
    main()
