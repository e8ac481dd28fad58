"""Golden vectors for the PPO notebook's client-side reward shaping, produced by EXECUTING the
notebook's own cells (Actor_Critic_PPO.ipynb code cells 6 and 7) on trajectories of the reference
engine.  Run here (needs /root/reference):  python tests/golden/make_shaping_golden.py

Also executes Policy_Gradients.ipynb's own calc_reward(state) (code cells 5-6) on the same episodes ->
tests/golden/shaping_pg_golden.npz (reward [E, T]).

Output tests/golden/shaping_golden.npz:
  obs      [E, T+1, 15]  observations s_0..s_T of E episodes (reference DroneGame, float64)
  n_steps  [E]           steps actually played (episode ends on done; later rows are padding)
  reward   [E, T]        calc_reward(s_{k+1}, prev_state = s_{k-1})['total']  (None for k = 0), -500 at time-out
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.modules.setdefault("pygame", types.ModuleType("pygame"))
sys.path.insert(0, "/root/reference")
from delivery_drone.game.game_engine import DroneGame            # noqa: E402
from delivery_drone.game.socket_client import DroneState          # noqa: E402
import rl_helpers.scalers as scalers                              # noqa: E402

KEYS = ("drone_x", "drone_y", "drone_vx", "drone_vy", "drone_angle", "drone_angular_vel", "drone_fuel", "platform_x",
        "platform_y", "distance_to_platform", "dx_to_platform", "dy_to_platform", "speed", "landed", "crashed")


def notebook_calc_reward():
    import math
    nb = json.load(open("/root/reference/Actor_Critic_PPO.ipynb"))
    ns = {"math": math, "np": np, "DroneState": DroneState}
    ns.update({k: getattr(scalers, k) for k in dir(scalers) if not k.startswith("_")})
    for idx in (6, 7):
        exec("".join(nb["cells"][idx]["source"]), ns)
    return ns["calc_reward"]


def notebook_calc_reward_pg():
    """Policy_Gradients.ipynb: calc_velocity_alignment (code cell 5) and calc_reward(state) (code cell 6)."""
    import math
    nb = json.load(open("/root/reference/Policy_Gradients.ipynb"))
    code = [c for c in nb["cells"] if c["cell_type"] == "code"]
    ns = {"math": math, "np": np, "DroneState": DroneState}
    ns.update({k: getattr(scalers, k) for k in dir(scalers) if not k.startswith("_")})
    for idx in (5, 6):
        exec("".join(code[idx]["source"]), ns)
    return ns["calc_reward"]


def main(E=48, T=120, max_steps=100):
    calc_reward = notebook_calc_reward()
    calc_reward_pg = notebook_calc_reward_pg()
    rew_pg = np.zeros((E, T))
    rng = np.random.default_rng(2024)
    obs = np.zeros((E, T + 1, 15))
    rew = np.zeros((E, T))
    n_steps = np.zeros(E, np.int32)
    for e in range(E):
        np.random.seed(1000 + e)
        g = DroneGame(render_mode=None, randomize_drone=True, randomize_platform=True)
        state = DroneState(**g.reset())
        prev = None
        # a mix of behaviours: bang-bang hover-ish controllers land sometimes, random ones crash / drift
        style = e % 4
        for k in range(T):
            obs[e, k] = [float(getattr(state, key)) for key in KEYS]
            if style == 0:
                a = rng.integers(0, 2, 3)
            elif style == 1:
                a = [int(state.drone_vy * 10 > 1.5), 0, 0]
            elif style == 2:
                a = [int(state.drone_vy * 10 > 1.0), int(state.dx_to_platform < -0.02 and rng.random() < 0.3),
                     int(state.dx_to_platform > 0.02 and rng.random() < 0.3)]
            else:
                a = [int(rng.random() < 0.45), int(rng.random() < 0.1), int(rng.random() < 0.1)]
            nxt, _, done, _ = g.step({"main_thrust": int(a[0]), "left_thrust": int(a[1]), "right_thrust": int(a[2])})
            nxt = DroneState(**nxt)
            r = calc_reward(nxt, prev_state=prev)["total"]
            r_pg = calc_reward_pg(nxt)["total"]
            if k + 1 >= max_steps:                          # Actor_Critic_PPO.ipynb c16:L89-93 / Policy_Gradients collect_episodes
                if not nxt.landed:
                    r -= 500
                    r_pg -= 500
                done = True
            rew[e, k] = r
            rew_pg[e, k] = r_pg
            prev, state = state, nxt                        # c16:L101-102
            n_steps[e] = k + 1
            if done:
                obs[e, k + 1] = [float(getattr(state, key)) for key in KEYS]
                break
    np.savez_compressed(os.path.join(HERE, "shaping_golden.npz"), obs=obs, reward=rew, n_steps=n_steps,
                        max_steps=np.int32(max_steps))
    np.savez_compressed(os.path.join(HERE, "shaping_pg_golden.npz"), reward=rew_pg)     # same episodes / observations
    landed = sum(obs[e, n_steps[e], 13] for e in range(E))
    print("episodes", E, "landed", landed, "lengths", n_steps.min(), n_steps.max(), "reward range", rew.min(), rew.max())


if __name__ == "__main__":
    main()
