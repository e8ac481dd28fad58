"""N2 -- the notebooks' client-side training reward (variant 'ppo': Actor_Critic_PPO / Actor_Critic_Basic
calc_reward(state, prev_state); variant 'pg': Policy_Gradients calc_reward(state)) as a fused epilogue of the rollout
kernels, against oracle/shaping_port.py (pinned to the notebooks' executed cells).

float64 instantiation: agreement to 1e-9 (same statement order, same roundings up to sin/cos ulps in
the underlying state).  float32: |a-b| <= 2e-3*max(1,|b|) -- the distance term multiplies a
difference of two normalised distances by 1000*(1+2*speed), amplifying fp32 rounding (1e-7) to
~1e-3 -- except where a branch condition of calc_reward sits within 1e-5 of its threshold."""
import importlib

import numpy as np
import pytest
import torch

from oracle.shaping_port import EpisodeShaper, DIST, SPEED, VX, VY, DX, DY

pytestmark = pytest.mark.gpu
dd = importlib.import_module("reinforcement-learning-101_b200")
DEV = "cuda:0"


def _expected(obs0, obs, done, max_steps, auto_reset, variant="ppo"):
    """obs0 [N,15] initial observation, obs [T,N,15] observation after each step, done [T,N] flags."""
    T, N = done.shape
    exp = np.zeros((T, N))
    valid = np.ones((T, N), bool)
    for i in range(N):
        sh = EpisodeShaper(max_steps, variant=variant)
        cur = obs0[i]
        for t in range(T):
            if auto_reset and done[t, i]:
                valid[t, i] = False                   # terminal observation is not in the stream
                sh.reset(); cur = obs[t, i]
                continue
            exp[t, i], _ = sh.step(cur, obs[t, i])
            cur = obs[t, i]
            if done[t, i]:
                exp[t + 1:, i] = 0.0                  # frozen after done: no step, no reward
                break
    return exp, valid


def _near_threshold(cur_prev_dist, o):
    """True where a branch condition of calc_reward is within 1e-5 of flipping."""
    d, s = o[..., DIST], o[..., SPEED]
    toward = (o[..., VX] * o[..., DX] + o[..., VY] * o[..., DY]) / np.maximum(d, 1e-9)
    # policy-gradient variant: sign of the velocity alignment, its speed limit
    align = -(o[..., VX] * o[..., DX] + o[..., VY] * o[..., DY])
    m = np.minimum.reduce([np.abs(s - 0.15), np.abs(s - 0.05), np.abs(toward - 0.1), np.abs(d - 0.065),
                           np.abs(d - 0.3), np.abs(o[..., DY]), np.abs(s - 0.1), np.abs(s - 0.6), np.abs(s - 0.4),
                           np.abs(align) * 1e2])
    return m < 1e-5


@pytest.mark.parametrize("variant", ["ppo", "pg"])
@pytest.mark.parametrize("dtype,policy", [(torch.float64, "bangbang"), (torch.float64, "random"), (torch.float32, "random")])
def test_shaped_reward_freeze_after_done(dtype, policy, variant):
    n, T, ms = 96, 130, 100
    e = dd.BatchedDroneEnv(n, device=DEV, seed=4, randomize_drone=True, randomize_platform=True, max_steps=ms,
                           auto_reset=False, dtype=dtype, shaping=variant)
    obs0 = e.reset().double().cpu().numpy().copy()
    obs = torch.empty(T, n, 15, dtype=dtype, device=DEV); don = torch.empty(T, n, dtype=torch.uint8, device=DEV)
    shp = torch.empty(T, n, dtype=dtype, device=DEV); rew = torch.empty(T, n, dtype=dtype, device=DEV)
    # two launches: prev_dist must carry over between them
    e.rollout(60, policy, obs_out=obs[:60], done_out=don[:60], shaped_out=shp[:60], reward_out=rew[:60])
    e.rollout(T - 60, policy, t0=60, obs_out=obs[60:], done_out=don[60:], shaped_out=shp[60:], reward_out=rew[60:])
    exp, valid = _expected(obs0, obs.double().cpu().numpy(), don.cpu().numpy(), ms, auto_reset=False, variant=variant)
    got = shp.double().cpu().numpy()
    assert valid.all()
    if dtype == torch.float64:
        np.testing.assert_allclose(got, exp, rtol=1e-9, atol=1e-9)
    else:
        bad = np.abs(got - exp) > 2e-3 * np.maximum(1.0, np.abs(exp))
        near = _near_threshold(None, obs.double().cpu().numpy())
        assert not (bad & ~near).any(), np.argwhere(bad & ~near)[:5]
        assert bad.sum() <= 5
    f = don.cpu().numpy()
    assert (f & 8).any()                                       # time-outs occur
    if policy == "random":
        assert (f & 4).any()                                   # ... and crashes
    if policy == "bangbang":
        assert (f & 2).any() and got.max() > (800 if variant == "ppo" else 500)   # landings: 800 (ppo) / 500 (pg) + 100 * fuel
    assert got[ms - 1][(f[ms - 1] & 8) > 0].max() < -490       # -500 on the time-out step (later rows: frozen, 0)
    assert (got[ms:] == 0).all()


@pytest.mark.parametrize("variant", ["ppo", "pg"])
def test_shaped_reward_with_auto_reset_and_fused_policy(golden_dir, variant):
    import os
    n, T, ms = 256, 120, 60
    kw = dict(seed=8, randomize_drone=True, randomize_platform=True, max_steps=ms, auto_reset=True, dtype=torch.float32,
              shaping=variant)
    e = dd.BatchedDroneEnv(n, device=DEV, **kw)
    obs0 = e.reset().double().cpu().numpy().copy()
    obs = torch.empty(T, n, 15, device=DEV); don = torch.empty(T, n, dtype=torch.uint8, device=DEV)
    shp = torch.empty(T, n, device=DEV)
    e.rollout(T, "random", obs_out=obs, done_out=don, shaped_out=shp)
    exp, valid = _expected(obs0, obs.double().cpu().numpy(), don.cpu().numpy(), ms, auto_reset=True, variant=variant)
    got = shp.double().cpu().numpy()
    bad = (np.abs(got - exp) > 2e-3 * np.maximum(1.0, np.abs(exp))) & valid
    near = _near_threshold(None, obs.double().cpu().numpy())
    assert not (bad & ~near).any(), np.argwhere(bad & ~near)[:5]
    assert (~valid).sum() > n                                   # several episodes per env ended
    # the fused policy kernel computes the same shaped reward as the plain rollout on the same actions
    d = np.load(os.path.join(golden_dir, "policy_v1.npz"))
    blob = dd.PolicyBlob({k: torch.from_numpy(d[k]) for k in d.files if k.startswith("network")}, device=DEV)
    a = dd.BatchedDroneEnv(n, device=DEV, **kw); a.reset()
    out = dd.policy_rollout(a, blob, T, sample=True, want="ars")
    b = dd.BatchedDroneEnv(n, device=DEV, **kw); b.reset()
    shp_b = torch.empty(T, n, device=DEV)
    b.rollout(T, "trace", actions=out["actions"], shaped_out=shp_b)
    assert torch.equal(out["shaped"], shp_b)
    assert torch.equal(a.get_state()["prev_dist"].isnan(), b.get_state()["prev_dist"].isnan())


@pytest.mark.parametrize("variant", ["ppo", "pg"])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_terminal_shaped_rewards_under_auto_reset(dtype, variant):
    """under auto-reset the observation stream holds the NEXT episode's first observation on a
    terminal step, so the tests above skip those steps -- but the +800 / -300 / -500 terms on exactly those steps are
    what PPO learns from.  Here a twin env replays the same actions step by step with want_final_obs=True; its
    terminal observation rows feed the notebook port, and EVERY step of the rollout is compared."""
    n, T, ms = 256, 150, 60
    kw = dict(seed=12, randomize_drone=True, randomize_platform=True, max_steps=ms, auto_reset=True, dtype=dtype, shaping=variant)
    a = dd.BatchedDroneEnv(n, device=DEV, **kw)
    obs0 = a.reset().double().cpu().numpy().copy()
    acts = a.random_actions(T)
    acts[:, : n // 2] &= 1                                    # half the envs only ever fire the main thruster: some land
    shp = torch.empty(T, n, dtype=dtype, device=DEV); don = torch.empty(T, n, dtype=torch.uint8, device=DEV)
    a.rollout(T, "trace", actions=acts, shaped_out=shp, done_out=don)
    b = dd.BatchedDroneEnv(n, device=DEV, want_final_obs=True, **kw); b.reset()
    exp = np.zeros((T, n))
    shapers = [EpisodeShaper(ms, variant=variant) for _ in range(n)]
    cur = obs0.copy()
    near = np.zeros((T, n), bool)
    for t in range(T):
        obs, rew, flags = b.step_raw(acts[t])
        obs = obs.double().cpu().numpy(); fin = b.final_obs.double().cpu().numpy(); fl = flags.cpu().numpy()
        assert np.array_equal(fl, don[t].cpu().numpy())
        for i in range(n):
            nxt = fin[i] if fl[i] else obs[i]                 # the state the notebook's calc_reward sees: the terminal one
            exp[t, i], _ = shapers[i].step(cur[i], nxt)
            if fl[i]:
                shapers[i].reset()
        nxt_all = np.where(fl[:, None] != 0, fin, obs)
        near[t] = _near_threshold(None, nxt_all)
        cur = obs                                             # after a reset: the new episode's first observation
    got = shp.double().cpu().numpy()
    f = don.cpu().numpy()
    term = f != 0
    assert term.sum() > 2 * n
    if dtype == torch.float64:
        np.testing.assert_allclose(got, exp, rtol=1e-9, atol=1e-9)
    else:
        bad = np.abs(got - exp) > 2e-3 * np.maximum(1.0, np.abs(exp))
        assert not (bad & ~near).any(), np.argwhere(bad & ~near)[:5]
        assert bad.sum() <= 5
    # the terminal terms themselves
    landed, crashed, trunc = (f & 2) > 0, (f & 4) > 0, (f & 8) > 0
    assert crashed.any() and trunc.any()
    assert (got[crashed] < -150).all()                                  # -200 / -300 (+ -500 when it is also the cap step)
    assert (got[trunc & ~landed] < -480).all()                          # -500 time-out
    if landed.any():
        assert (got[landed] > (790 if variant == "ppo" else 490)).all()  # +800 / +500 + 100 * fuel, never the -500
