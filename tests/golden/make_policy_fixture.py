"""Records the trained policy checkpoint of the reference as a small fixture (weights are data,
not code): /root/reference/models/actor-critic-ppo/drone_policy_v1.pth -> tests/golden/policy_v1.npz,
plus the eager-torch fp32 outputs of the notebook's network (Actor_Critic_PPO.ipynb c11:L5-17) on a
fixed observation batch, so the fused tensor-core path can be checked on a box without the reference.

Run here (needs /root/reference):  python tests/golden/make_policy_fixture.py
"""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/models/actor-critic-ppo/drone_policy_v1.pth"
SRC_CRITIC = "/root/reference/models/actor-critic-ppo/drone_critic_v1.pth"     # DroneTeacherBoi, same trunk, 1 output


def fixed_obs(n=512, seed=1234):
    """Plausible normalised observations (the 15 policy inputs) from a numpy Generator."""
    r = np.random.default_rng(seed)
    x, y = r.uniform(0.05, 0.95, n), r.uniform(0.0, 0.95, n)
    px, py = r.uniform(0.125, 0.875, n), r.uniform(0.17, 0.92, n)
    vx, vy = r.normal(0, 0.3, n), r.normal(0.4, 0.4, n)
    ang, angv = r.normal(0, 0.2, n), r.normal(0, 0.15, n)
    fuel = r.uniform(0.3, 1.0, n)
    dx, dy = px - x, py - y
    dist = np.sqrt((dx * 800) ** 2 + (dy * 600) ** 2) / 800
    speed = np.sqrt(vx ** 2 + vy ** 2)
    z = np.zeros(n)
    return np.stack([x, y, vx, vy, ang, angv, fuel, px, py, dist, dx, dy, speed, z, z], 1).astype(np.float32)


def main():
    sd = torch.load(SRC, weights_only=True, map_location="cpu")
    net = torch.nn.Sequential(
        torch.nn.Linear(15, 128), torch.nn.LayerNorm(128), torch.nn.ReLU(),
        torch.nn.Linear(128, 128), torch.nn.LayerNorm(128), torch.nn.ReLU(),
        torch.nn.Linear(128, 64), torch.nn.LayerNorm(64), torch.nn.ReLU(),
        torch.nn.Linear(64, 3), torch.nn.Sigmoid())
    net.load_state_dict({k.replace("network.", ""): v for k, v in sd.items()})
    obs = fixed_obs()
    with torch.no_grad():
        probs = net(torch.from_numpy(obs)).numpy()
        logits = net[:-1](torch.from_numpy(obs)).numpy()
    out = {k: v.numpy().astype(np.float32) for k, v in sd.items()}
    out["obs"], out["probs"], out["logits"] = obs, probs, logits
    np.savez_compressed(os.path.join(HERE, "policy_v1.npz"), **out)
    print("wrote policy_v1.npz;", sum(v.size for k, v in out.items() if k.startswith("network")), "parameters;",
          "logit range", logits.min(), logits.max())
    # the critic of the same run: values on the same observation batch
    sdc = torch.load(SRC_CRITIC, weights_only=True, map_location="cpu")
    crit = torch.nn.Sequential(
        torch.nn.Linear(15, 128), torch.nn.LayerNorm(128), torch.nn.ReLU(),
        torch.nn.Linear(128, 128), torch.nn.LayerNorm(128), torch.nn.ReLU(),
        torch.nn.Linear(128, 64), torch.nn.LayerNorm(64), torch.nn.ReLU(),
        torch.nn.Linear(64, 1))
    crit.load_state_dict({k.replace("network.", ""): v for k, v in sdc.items()})
    with torch.no_grad():
        values = crit(torch.from_numpy(obs)).squeeze(-1).numpy()
    outc = {k: v.numpy().astype(np.float32) for k, v in sdc.items()}
    outc["obs"], outc["values"] = obs, values
    np.savez_compressed(os.path.join(HERE, "critic_v1.npz"), **outc)
    print("wrote critic_v1.npz;", sum(v.size for k, v in outc.items() if k.startswith("network")), "parameters;",
          "value range", values.min(), values.max())


if __name__ == "__main__":
    main()
