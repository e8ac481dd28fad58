"""Parity of the sm_100a kernels (through the C ABI) against the float64 oracle and the golden
vectors recorded from the unmodified reference.  GPU only (``-m gpu``).

Tolerances (north_star):
  * float64 instantiation: flags / step counters bit-exact, continuous values to 1e-11
    (device sincos vs libm: <= 1-2 ulp, amplified by 250 steps of accumulation).
  * float32 instantiation: ``|a - b| <= 1e-5 * max(|a|, 1)`` on normalised observations and on
    rewards (rtol 1e-5 + atol 1e-5; SURVEY.md 7 hard-part 1), flags / counters bit-exact EXCEPT
    for envs whose float64 state lies within ``BAND`` of a termination threshold on the step
    where the decision differs; those are listed, bounded in number, and compared only up to
    that step.  BAND is 2e-4 px (or deg, px/step): north_star's literal 1e-6 band is smaller
    than the accumulated fp32 position error (~1.7e-3 px after 250 steps) and cannot be met by
    any fp32 state (the float64 instantiation is the path that meets it: bit-exact); observed on
    B200: one flip per run, with float64 margins 1.9e-5 / 4.1e-5 px -- a flip rate of a few 1e-5
    per episode.  The same bound holds for the HOST instantiation of the same source
    (tests/test_host_twin_cpu.py).
"""
import importlib
import json
import os

import numpy as np
import pytest
import torch

from corpus import N_CORPUS, N_TRAJ, T_CORPUS, corpus_actions, corpus_spawns
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu


def _eq(a, b):
    """torch.equal that treats NaN == NaN (prev_dist uses NaN for 'no previous state')."""
    if a.is_floating_point():
        a, b = torch.nan_to_num(a, nan=-12345.0), torch.nan_to_num(b, nan=-12345.0)
    return torch.equal(a, b)

dd = importlib.import_module("reinforcement-learning-101_b200")
nv = dd.native

BAND = 2e-4
OBS_NAMES = ("drone_x", "drone_y", "drone_vx", "drone_vy", "drone_angle", "drone_angular_vel", "drone_fuel",
             "platform_x", "platform_y", "distance_to_platform", "dx_to_platform", "dy_to_platform", "speed",
             "landed", "crashed")


def _env(n, dtype=torch.float64, **kw):
    kw.setdefault("randomize_drone", False)
    kw.setdefault("randomize_platform", False)
    kw.setdefault("auto_reset", False)
    return dd.BatchedDroneEnv(n, device="cuda:0", dtype=dtype, **kw)


def _t(a, dtype=None):
    return torch.as_tensor(np.ascontiguousarray(a), device="cuda:0") if dtype is None else \
        torch.as_tensor(np.ascontiguousarray(a), device="cuda:0").to(dtype)


def _close32(a, b):
    """north_star fp32 tolerance."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b) <= 1e-5 * np.maximum(np.abs(b), 1.0)


def _threshold_margin(o):
    """Smallest distance of the oracle's float64 state to any termination threshold
    (game_engine.py:224-279), per env, in raw units."""
    rad = np.radians(o.angle)
    bx, by = o.x - 10.0 * np.sin(rad), o.y + 10.0 * np.cos(rad)
    speed = np.sqrt(o.vx ** 2 + o.vy ** 2)
    m = [np.abs(speed - 3.0), np.abs(np.abs(o.angle) - 20.0),
         np.abs(bx - (o.px - 50)), np.abs(bx - (o.px + 50)), np.abs(by - (o.py - 10)), np.abs(by - (o.py + 10)),
         np.abs(o.y - 550.0), np.abs(o.x + 50.0), np.abs(o.x - 850.0), np.abs(o.y + 50.0)]
    return np.min(np.stack(m), axis=0)


# =================================================================================================
# config 2: 4,096 envs, fixed action trace, 250 steps, freeze-after-done
# =================================================================================================
def test_corpus_f64_vs_oracle_and_golden(golden_dir):
    gs = np.load(os.path.join(golden_dir, "corpus_summary.npz"))
    gt = np.load(os.path.join(golden_dir, "corpus_traj.npz"))
    sx, sy, spx, spy = corpus_spawns()
    A = corpus_actions()
    o = co.OracleBatch(N_CORPUS, randomize_drone=False, randomize_platform=False)
    o.inject(sx, sy, spx, spy)
    e = _env(N_CORPUS)
    obs0 = e.inject(_t(sx), _t(sy), _t(spx), _t(spy)).cpu().numpy()
    np.testing.assert_allclose(obs0, o.write_obs_into(np.zeros((N_CORPUS, 15))), rtol=1e-15, atol=0)
    Ad = _t(A)
    done_step = np.zeros(N_CORPUS, np.int16)
    for t in range(T_CORPUS):
        oo, orr, od = o.step(A[t])
        obs, rew, done, info = e.step(Ad[t])
        obs, rew, fl = obs.cpu().numpy(), rew.cpu().numpy(), info["flags"].cpu().numpy()
        assert np.array_equal(fl, od), f"flags differ at step {t}"
        assert np.array_equal(done.cpu().numpy(), (od & co.DONE) > 0)
        assert np.array_equal(e.steps.cpu().numpy(), o.steps), f"step counters differ at step {t}"
        np.testing.assert_allclose(obs, oo, rtol=1e-11, atol=1e-12)
        np.testing.assert_allclose(rew, orr, rtol=1e-11, atol=1e-12)
        # golden vectors recorded from the unmodified reference
        np.testing.assert_allclose(obs.sum(0), gs["obs_sum"][t], rtol=1e-11, atol=1e-8)
        np.testing.assert_allclose(obs[:N_TRAJ], gt["obs"][t], rtol=1e-11, atol=1e-12)
        np.testing.assert_allclose(rew[:N_TRAJ], gt["reward"][t], rtol=1e-11, atol=1e-12)
        assert int(((fl & co.DONE) > 0).sum()) == gs["done_cnt"][t]
        d = (fl & co.DONE) > 0
        newly = d & (done_step == 0)
        done_step[newly] = e.steps.cpu().numpy()[newly]
    assert np.array_equal(done_step, gs["done_step"])
    st = e.get_state()
    assert np.array_equal(st["flags"].cpu().numpy() & 7, gs["flags"])
    np.testing.assert_allclose(st["total_reward"].cpu().numpy(), gs["total"], rtol=1e-11, atol=1e-11)
    fin = np.stack([st[k].cpu().numpy() for k in ("x", "y", "vx", "vy", "angle", "angular_velocity", "fuel")], 1)
    np.testing.assert_allclose(fin, gs["final"], rtol=1e-11, atol=1e-11)


def test_corpus_f32_vs_oracle():
    sx, sy, spx, spy = corpus_spawns()
    A = corpus_actions()
    o = co.OracleBatch(N_CORPUS, randomize_drone=False, randomize_platform=False)
    o.inject(sx, sy, spx, spy)
    e = _env(N_CORPUS, dtype=torch.float32)
    e.inject(_t(sx, torch.float32), _t(sy, torch.float32), _t(spx, torch.float32), _t(spy, torch.float32))
    Ad = _t(A)
    valid = np.ones(N_CORPUS, bool)          # envs still under comparison
    flips = []
    worst = 0.0
    for t in range(T_CORPUS):
        margin = None
        oo, orr, od = o.step(A[t])
        obs, rew, done, info = e.step(Ad[t])
        fl = info["flags"].cpu().numpy()
        bad = valid & (fl != od)
        if bad.any():
            margin = _threshold_margin(o)
            for i in np.nonzero(bad)[0]:
                assert margin[i] < BAND, f"env {i} step {t}: flags {fl[i]:#x} vs {od[i]:#x}, margin {margin[i]:.3g}"
                flips.append((int(i), t, float(margin[i])))
            valid &= ~bad
        assert np.array_equal(e.steps.cpu().numpy()[valid], o.steps[valid])
        obs, rew = obs.cpu().numpy(), rew.cpu().numpy()
        ok = _close32(obs, oo)[valid]
        assert ok.all(), f"step {t}: obs outside fp32 tolerance, max abs err {np.abs(obs - oo)[valid].max():.3g}"
        assert _close32(rew, orr)[valid].all(), f"step {t}: reward outside fp32 tolerance"
        worst = max(worst, float(np.abs(obs - oo)[valid].max()))
    assert len(flips) <= 2, flips            # ~3e-5 per episode expected (3,650 episodes in the corpus)
    print(f"fp32 corpus: {len(flips)} threshold flips {flips}, worst normalised-obs abs error {worst:.3g}")


# =================================================================================================
# KATs recorded from the reference (fixed spawn)
# =================================================================================================
KAT_ACTION = {"KAT1_no_thrust": 0, "KAT2_main": 1, "KAT3_all": 7, "KAT4_right": 4, "KAT5_main_right": 5}


@pytest.mark.parametrize("name", sorted(KAT_ACTION) + ["KAT6_bangbang"])
def test_kat_f64(golden_dir, name):
    ref = json.load(open(os.path.join(golden_dir, "kat.json")))[name]
    T = ref["steps"]
    e = _env(3)
    e.reset()
    rew = torch.zeros(T + 2, 3, dtype=torch.float64, device="cuda:0")
    don = torch.zeros(T + 2, 3, dtype=torch.uint8, device="cuda:0")
    if name == "KAT6_bangbang":
        e.rollout(T + 2, policy="bangbang", reward_out=rew, done_out=don)
    else:
        a = torch.full((T + 2, 3), KAT_ACTION[name], dtype=torch.uint8, device="cuda:0")
        e.rollout(T + 2, policy="trace", actions=a, reward_out=rew, done_out=don)
    rew, don = rew.cpu().numpy()[:, 1], don.cpu().numpy()[:, 1]
    assert not don[:T - 1].any() and (don[T - 1] & co.DONE)
    assert bool(don[T - 1] & co.LANDED) == ref["landed"] and bool(don[T - 1] & co.CRASHED) == ref["crashed"]
    tol = dict(rel=1e-12, abs=1e-11)
    assert rew[T - 1] == pytest.approx(ref["last_reward"], **tol)
    assert rew[T] == 0.0 and rew[T + 1] == 0.0 and (don[T] & co.DONE)      # frozen after done
    st = {k: v.cpu().numpy()[1] for k, v in e.get_state().items()}
    assert st["steps"] == ref["steps"]
    for k, rk in (("x", "x"), ("y", "y"), ("vx", "vx"), ("vy", "vy"), ("angle", "angle"),
                  ("angular_velocity", "angvel"), ("fuel", "fuel"), ("total_reward", "total_reward")):
        assert float(st[k]) == pytest.approx(ref[rk], **tol), k
    obs = e.observe().cpu().numpy()[1]
    for j, k in enumerate(OBS_NAMES):
        assert obs[j] == pytest.approx(float(ref["final_state"][k]), **tol), k
    for h, r in zip(ref["head"], rew[:3]):
        assert r == pytest.approx(h["r"], **tol)


# =================================================================================================
# auto-reset, Philox spawns, truncation, statistics -- float64 is bit-exact with the oracle
# =================================================================================================
@pytest.mark.parametrize("n", [1, 255, 257, 5000])
def test_autoreset_random_policy_f64_exact(n):
    kw = dict(seed=11, randomize_drone=True, randomize_platform=True, max_steps=60, auto_reset=True, env_id_base=1000)
    o = co.OracleBatch(n, **kw)
    o.reset()
    e = dd.BatchedDroneEnv(n, device="cuda:0", dtype=torch.float64, want_final_obs=True, **kw)
    obs0 = e.reset().cpu().numpy()
    assert np.array_equal(e.pos_vel[:, 0].cpu().numpy(), o.x) and np.array_equal(e.platform[:, 1].cpu().numpy(), o.py)
    np.testing.assert_allclose(obs0, o.write_obs_into(np.zeros((n, 15))), rtol=1e-15)
    T = 150
    A = co.random_actions(11, 1000, 0, T, n)
    assert np.array_equal(e.random_actions(T).cpu().numpy(), A)
    Ad = _t(A)
    for t in range(T):
        oo, orr, od, of = o.step(A[t], want_final=True)
        obs, rew, done, info = e.step(Ad[t])
        fl = info["flags"].cpu().numpy()
        assert np.array_equal(fl, od), t
        np.testing.assert_allclose(obs.cpu().numpy(), oo, rtol=1e-11, atol=1e-12)
        np.testing.assert_allclose(rew.cpu().numpy(), orr, rtol=1e-11, atol=1e-12)
        d = (fl & co.DONE) > 0
        np.testing.assert_allclose(info["final_obs"].cpu().numpy()[d], of[d], rtol=1e-11, atol=1e-12)
        assert np.array_equal(e.steps.cpu().numpy(), o.steps)
        assert np.array_equal(e.episode.cpu().numpy().astype(np.uint32), o.episode)
    s = e.stats()
    assert (s["episodes"], s["landed"], s["crashed"], s["truncated"]) == \
        (o.stats.episodes, o.stats.landed, o.stats.crashed, o.stats.truncated)
    assert s["episodes"] > 0 and s["env_steps"] == n * T
    assert s["sum_length"] == o.stats.sum_length
    assert s["sum_return"] == pytest.approx(o.stats.sum_return, rel=1e-9, abs=1e-5 * max(1, s["episodes"]))


@pytest.mark.parametrize("policy", ["random", "bangbang", "trace"])
def test_rollout_kernel_matches_oracle_rollout(policy):
    n, T = 3000, 130
    kw = dict(seed=5, randomize_drone=True, randomize_platform=True, max_steps=100, auto_reset=True, env_id_base=77)
    o = co.OracleBatch(n, **kw)
    o.reset()
    e = dd.BatchedDroneEnv(n, device="cuda:0", dtype=torch.float64, **kw)
    e.reset()
    A = co.random_actions(99, 0, 0, T, n) if policy == "trace" else None
    pol = {"random": co.POL_RANDOM, "bangbang": co.POL_BANGBANG, "trace": co.POL_TRACE}[policy]
    orew, odon, ostats = o.rollout(T, policy=pol, actions=A, t0=7, record=True)
    rew = torch.empty(T, n, dtype=torch.float64, device="cuda:0")
    don = torch.empty(T, n, dtype=torch.uint8, device="cuda:0")
    obs = torch.empty(T, n, 15, dtype=torch.float64, device="cuda:0")
    e.rollout(T, policy=policy, actions=None if A is None else _t(A), t0=7, reward_out=rew, done_out=don, obs_out=obs)
    assert np.array_equal(don.cpu().numpy(), odon)
    np.testing.assert_allclose(rew.cpu().numpy(), orew, rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(e.observe().cpu().numpy(), o.write_obs_into(np.zeros((n, 15))), rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(obs[-1].cpu().numpy(), e.observe().cpu().numpy(), rtol=0, atol=0)
    s = e.stats()
    assert (s["episodes"], s["landed"], s["crashed"], s["truncated"], s["sum_length"]) == \
        (ostats.episodes, ostats.landed, ostats.crashed, ostats.truncated, ostats.sum_length)
    assert s["env_steps"] == n * T


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_rollout_equals_stepping(dtype):
    """T-steps-per-launch and one-step-per-launch kernels run the same arithmetic: bit-identical."""
    n, T = 4097, 80
    kw = dict(seed=3, randomize_drone=True, randomize_platform=True, max_steps=50, auto_reset=True, dtype=dtype)
    a = dd.BatchedDroneEnv(n, device="cuda:0", **kw); a.reset()
    b = dd.BatchedDroneEnv(n, device="cuda:0", **kw); b.reset()
    A = a.random_actions(T)
    rew = torch.empty(T, n, dtype=dtype, device="cuda:0")
    don = torch.empty(T, n, dtype=torch.uint8, device="cuda:0")
    obs = torch.empty(T, n, 16, dtype=dtype, device="cuda:0")
    b16 = dd.BatchedDroneEnv(n, device="cuda:0", obs_stride=16, **kw); b16.reset()
    b16.rollout(T, policy="random", reward_out=rew, done_out=don, obs_out=obs)
    b.rollout(T, policy="trace", actions=A)
    for t in range(T):
        o_, r_, f_ = a.step_raw(A[t])
        assert torch.equal(r_, rew[t]) and torch.equal(f_, don[t]) and torch.equal(o_, obs[t, :, :15])
    assert torch.equal(obs[-1, :, 15], a.steps.to(dtype))          # 16th key: steps
    for k, v in a.get_state().items():
        assert _eq(v, b.get_state()[k]), k
        assert _eq(v, b16.get_state()[k]), k
    assert a.stats() == b.stats() == b16.stats()


def test_sharding_invariance():
    """Global env ids key the spawns: 2 shards of N/2 == 1 env of N (any GPU count gives the same
    trajectories)."""
    n, T = 2048, 100
    kw = dict(seed=9, randomize_drone=True, randomize_platform=True, max_steps=64, auto_reset=True, dtype=torch.float32)
    whole = dd.BatchedDroneEnv(n, device="cuda:0", **kw); whole.reset(); whole.rollout(T, "random")
    lo = dd.BatchedDroneEnv(n // 2, device="cuda:0", env_id_base=0, **kw); lo.reset(); lo.rollout(T, "random")
    hi = dd.BatchedDroneEnv(n // 2, device="cuda:0", env_id_base=n // 2, **kw); hi.reset(); hi.rollout(T, "random")
    sw, sl, sh = whole.get_state(), lo.get_state(), hi.get_state()
    for k in sw:
        assert _eq(sw[k], torch.cat([sl[k], sh[k]])), k
    a, b, c = whole.stats(), lo.stats(), hi.stats()
    for k in ("episodes", "landed", "crashed", "truncated", "sum_length", "env_steps", "sum_return"):
        assert a[k] == b[k] + c[k], k


def test_spawn_ranges_uniformity_f32():
    n = 400_000
    e = dd.BatchedDroneEnv(n, device="cuda:0", seed=7, randomize_drone=True, randomize_platform=True)
    obs = e.reset()
    o = co.OracleBatch(n, seed=7, randomize_drone=True, randomize_platform=True)
    o.reset()
    st = e.get_state()
    for k, ok, lo, hi in (("x", o.x, 100, 700), ("y", o.y, 50, 250), ("platform_x", o.px, 100, 699),
                          ("platform_y", o.py, 100, 549)):
        v = st[k].cpu().numpy().astype(np.float64)
        assert np.array_equal(v, ok) and v.min() == lo and v.max() == hi
    assert torch.all(st["fuel"] == 1000) and torch.all(st["episode"] == 1) and torch.all(obs[:, 6] == 1)
    f = dd.BatchedDroneEnv(4, device="cuda:0", randomize_drone=False, randomize_platform=False)
    f.reset()
    assert f.pos_vel[0, :2].tolist() == [400.0, 100.0] and f.platform[0].tolist() == [400.0, 500.0]


# =================================================================================================
# edge cases the reference defines
# =================================================================================================
def test_fuel_gating_order_and_skip():
    e = _env(3)
    e.reset()
    e.set_state({"fuel": torch.tensor([2.0, 1.0, 2.0], dtype=torch.float64)})
    before = e.get_state()
    obs, rew, done, info = e.step(torch.tensor([7, 7, 7 | nv.ACT_SKIP], dtype=torch.uint8, device="cuda:0"))
    st = e.get_state()
    assert st["fuel"].tolist() == [0.0, 0.0, 2.0] and st["angular_velocity"].tolist() == [0.0, 0.0, 0.0]
    fl = info["flags"].cpu().numpy()
    assert (fl[0] & co.CAUSE_MASK) == co.CAUSE_FUEL and (fl[1] & co.CAUSE_MASK) == co.CAUSE_FUEL and fl[2] == 0
    assert rew.tolist()[:2] == [pytest.approx(-50.1), pytest.approx(-50.1)] and rew[2].item() == 0.0
    for k in before:                                          # skipped env untouched
        if k != "prev_dist":                                  # NaN marker
            assert before[k][2].item() == st[k][2].item(), k
    assert st["steps"].tolist() == [1, 1, 0]


def test_actions_n3_truthiness_and_masked_reset():
    e = _env(4, randomize_platform=True, seed=2)
    e.reset()
    a3 = torch.tensor([[0, 0, 0], [2.5, 0, 0], [0, -1, 0], [0, 0, 1]], dtype=torch.float32, device="cuda:0")
    assert e.pack_actions(a3).tolist() == [0, 1, 2, 4]
    assert e.pack_actions(a3.ne(0)).tolist() == [0, 1, 2, 4]
    assert e.pack_actions(a3.ne(0).to(torch.uint8) * 200).tolist() == [0, 1, 2, 4]
    for _ in range(5):
        e.step(a3)
    before = e.get_state()
    obs = e.reset(mask=torch.tensor([1, 0, 0, 1], device="cuda:0"))
    st = e.get_state()
    assert st["steps"].tolist() == [0, 5, 5, 0] and st["episode"].tolist() == [2, 1, 1, 2]
    assert torch.equal(st["x"][1:3], before["x"][1:3]) and torch.equal(st["vy"][1:3], before["vy"][1:3])
    assert obs[0, 3].item() == 0.0 and obs[1, 3].item() == pytest.approx(before["vy"][1].item() / 10)


def test_empty_and_errors():
    e = _env(0)
    e.reset()
    o, r, d, _ = e.step(torch.zeros(0, dtype=torch.uint8, device="cuda:0"))
    assert o.shape == (0, 15) and r.shape == (0,) and d.shape == (0,)
    e.rollout(5, "random")
    assert e.stats()["episodes"] == 0
    e2 = _env(8)
    with pytest.raises(RuntimeError):
        e2.step(torch.zeros(8, dtype=torch.uint8, device="cuda:0"))
    e2.reset()
    with pytest.raises(ValueError):
        e2.step(torch.zeros(7, dtype=torch.uint8, device="cuda:0"))
    with pytest.raises(ValueError):
        e2.rollout(4, "trace")
    # C-ABI argument errors come back as codes, never as crashes
    import ctypes as C
    L = nv.lib()
    assert L.dd_step(None, None, None, None, None, 15, None, None, None, None, 8, None) == -1
    bad = nv.DDState(e2.pos_vel.data_ptr() + 4, e2.att_fuel.data_ptr(), e2.platform.data_ptr(), e2.steps.data_ptr(),
                     e2.episode.data_ptr(), e2.flags.data_ptr(), nv.F32)
    assert L.dd_step(C.byref(bad), C.byref(e2.params), C.byref(e2._cfg), e2._packed.data_ptr(), None, 15, None, None,
                     None, None, 8, None) == -4
    bad.dtype = 7
    assert L.dd_reset(C.byref(bad), C.byref(e2.params), C.byref(e2._cfg), None, None, 15, 8, None) == -3
    assert L.dd_step(C.byref(e2._state), C.byref(e2.params), C.byref(e2._cfg), e2._packed.data_ptr(),
                     e2.obs.data_ptr(), 17, None, None, None, None, 8, None) == -2
    assert b"null" in L.dd_error_string(-1)


def test_angle_wrap_and_closed_platform_box():
    """physics.py:35-39 wrap at +-180 (strict), platform.py:74 closed intervals, speed <= 3.0 and
    |angle| <= 20.0 inclusive (game_engine.py:232-237)."""
    e = _env(4)
    e.reset()
    z = torch.zeros(4, dtype=torch.float64)
    # env0: angle 179 + 2 -> 181 -> -179 ; env1: angle exactly 180 stays
    e.set_state({"angle": torch.tensor([179.0, 178.0, 0.0, 0.0], dtype=torch.float64),
                 "angular_velocity": torch.tensor([2.0, 2.0, 0.0, 0.0], dtype=torch.float64)})
    e.step(torch.zeros(4, dtype=torch.uint8, device="cuda:0"))
    assert e.get_state()["angle"].tolist()[:2] == [-179.0, 180.0]
    # bottom centre exactly on the platform's top-left corner with vy chosen so post-update values are exact
    e = _env(2)
    e.reset()
    # after update: vy' = (vy + 0.3) * 0.99, y' = y + vy'.  Pick vy = -0.3 -> vy' = 0 (allow -0.0), y' = y.
    e.set_state({"x": torch.tensor([350.0, 349.99], dtype=torch.float64), "y": torch.tensor([480.0, 480.0], dtype=torch.float64),
                 "vy": torch.tensor([-0.3, -0.3], dtype=torch.float64)})
    _, rew, done, info = e.step(torch.zeros(2, dtype=torch.uint8, device="cuda:0"))
    assert info["landed"].tolist() == [True, False] and rew[0].item() == pytest.approx(99.9)


# =================================================================================================
# K4 / N3: moments, normalisation, GAE
# =================================================================================================
@pytest.mark.parametrize("n", [1, 5, 1023, 1_000_003])
def test_moments_and_normalize(n):
    g = torch.Generator(device="cuda:0").manual_seed(n)
    x = (torch.randn(n + 1, device="cuda:0", generator=g) * 3 + 0.5)[1:].contiguous()      # 4-byte aligned only
    m = dd.advantage_moments(x).cpu().numpy()
    ref = co.moments(x.cpu().numpy())
    assert m[0] == n
    np.testing.assert_allclose(m[1:], ref[1:], rtol=1e-12, atol=1e-9)
    if n > 1:
        y = dd.normalize_advantages(x, reduce=False)
        t = (x - x.mean()) / (x.std() + 1e-8)                  # the notebook's expression, fp32 torch
        np.testing.assert_allclose(y.cpu().numpy(), t.cpu().numpy(), rtol=2e-5, atol=2e-5)
        x64 = x.double()
        t64 = (x64 - x64.mean()) / (x64.std() + 1e-8)
        np.testing.assert_allclose(y.cpu().numpy(), t64.cpu().numpy(), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("T,n", [(1, 1), (250, 1000), (33, 4099), (251, 4098), (19, 65536), (20, 6)])
def test_gae_bit_exact(T, n):
    """n covers every vector width of the scan (n % 4 == 0, n % 2 == 0, odd) and T the chunk tails."""
    rng = np.random.default_rng(T * 1000 + n)
    r = rng.normal(size=(T, n)).astype(np.float32)
    v = rng.normal(size=(T + 1, n)).astype(np.float32)
    d = (rng.random((T, n)) < 0.02).astype(np.uint8)
    adv, ret = dd.gae(_t(r), _t(v), _t(d), 0.99, 0.95, want_returns=True)
    ref = co.gae(r, v, d, 0.99, 0.95)
    assert np.array_equal(adv.cpu().numpy(), ref)
    assert np.array_equal(ret.cpu().numpy(), ref + v[:-1])
    # the moments accumulated in the same pass == dd_moments over the written buffer; unaligned views fall back
    m = torch.zeros(3, dtype=torch.float64, device="cuda:0")
    adv2 = dd.gae(_t(r), _t(v), _t(d), 0.99, 0.95, moments=m)
    assert torch.equal(adv2, adv)
    m2 = dd.advantage_moments(adv).cpu().numpy()
    assert m[0].item() == T * n
    np.testing.assert_allclose(m.cpu().numpy()[1:], m2[1:], rtol=1e-11, atol=1e-8)
    if n > 1:
        y1 = dd.normalize_advantages(adv, reduce=False)
        y2 = dd.normalize_advantages(adv, reduce=False, moments=m)
        np.testing.assert_allclose(y1.cpu().numpy(), y2.cpu().numpy(), rtol=1e-6, atol=1e-6)
    big = torch.zeros((T + 1) * n + 1, device="cuda:0")
    rv = big[1:T * n + 1].view(T, n); rv.copy_(_t(r))                  # 4-byte aligned only
    assert np.array_equal(dd.gae(rv, _t(v), _t(d), 0.99, 0.95).cpu().numpy(), ref)
    # and the notebook's own loop (Actor_Critic_PPO.ipynb c15:L49-53) in eager torch, one env
    i = n // 2
    rt, vt, dt = torch.tensor(r[:, i]), torch.tensor(v[:, i]), torch.tensor(d[:, i], dtype=torch.float32)
    gae_, out = torch.tensor(0.0), torch.zeros(T)
    for t in reversed(range(T)):
        mask = 1.0 - dt[t]
        delta = rt[t] + 0.99 * vt[t + 1] * mask - vt[t]
        gae_ = delta + 0.99 * 0.95 * mask * gae_
        out[t] = gae_
    assert np.array_equal(adv.cpu().numpy()[:, i], out.numpy())


# =================================================================================================
# BASELINE.json full size (1M envs): size-independent properties
# =================================================================================================
@pytest.mark.parametrize("T,n", [(1, 1), (37, 130), (250, 4099)])
def test_discounted_returns_bit_exact(T, n):
    """compute_returns of Policy_Gradients.ipynb (python-float loop G = r + gamma * G over reversed(rewards)),
    per episode: float64 accumulation, fp32 result -- bit for bit."""
    g = np.random.default_rng(T * 1000 + n)
    rew = g.normal(0, 3, (T, n)).astype(np.float32)
    done = (g.random((T, n)) < 0.02).astype(np.uint8)
    exp = np.zeros((T, n), np.float32)
    for i in range(min(n, 64)):                                  # the notebook's loop, episode by episode
        G = 0.0
        for t in reversed(range(T)):
            if done[t, i]:
                G = 0.0
            G = float(rew[t, i]) + 0.99 * G
            exp[t, i] = np.float32(G)
    got = dd.discounted_returns(_t(rew), _t(done), gamma=0.99).cpu().numpy()
    assert np.array_equal(got[:, :min(n, 64)], exp[:, :min(n, 64)])
    # vectorised float64 check of every column, and the no-dones form
    G = np.zeros(n); full = np.zeros((T, n), np.float32)
    for t in reversed(range(T)):
        G = np.where(done[t] != 0, 0.0, G)
        G = rew[t].astype(np.float64) + 0.99 * G
        full[t] = G.astype(np.float32)
    assert np.array_equal(got, full)
    nod = dd.discounted_returns(_t(rew), None, gamma=0.9).cpu().numpy()
    G = np.zeros(n)
    for t in reversed(range(T)):
        G = rew[t].astype(np.float64) + 0.9 * G
        assert np.array_equal(nod[t], G.astype(np.float32))


def test_full_size_properties():
    n, T = 1 << 20, 300
    kw = dict(seed=0, randomize_drone=True, randomize_platform=True, max_steps=250, auto_reset=True, dtype=torch.float32)
    a = dd.BatchedDroneEnv(n, device="cuda:0", **kw); a.reset(); a.rollout(T, "random")
    b = dd.BatchedDroneEnv(n, device="cuda:0", **kw); b.reset()
    A = b.random_actions(T)
    for t in range(T):
        b.step_raw(A[t], want_obs=(t % 50 == 0))
    sa, sb = a.get_state(), b.get_state()
    for k in sa:
        assert _eq(sa[k], sb[k]), k                      # idempotent / launch-shape independent
    s = a.stats()
    assert s == b.stats()
    assert s["env_steps"] == n * T
    assert s["episodes"] == s["landed"] + s["crashed"] + s["truncated"]
    assert s["episodes"] == int(sa["episode"].sum().item()) - n  # one reset() + one per finished episode
    assert 0.0 < s["landing_rate"] < 0.2 and 50 < s["mean_length"] <= 250
    # every live env is inside the arena and has fuel (else it would have terminated)
    assert torch.all(sa["steps"] < 250) and torch.all(sa["fuel"] > 0) and torch.all(sa["y"] <= 550)
    # C oracle on a 2,048-env slice of the same job (global ids 1,000,000..): float32 vs float64
    base, m = 1_000_000, 2048
    o = co.OracleBatch(m, seed=0, randomize_drone=True, randomize_platform=True, max_steps=250, auto_reset=True,
                       env_id_base=base)
    o.reset()
    _, _, ost = o.rollout(T, policy=co.POL_RANDOM)
    c = dd.BatchedDroneEnv(m, device="cuda:0", env_id_base=base, **kw); c.reset(); c.rollout(T, "random")
    for k in sa:
        assert _eq(c.get_state()[k], sa[k][base:base + m]), k
    sc = c.stats()
    # fp32 threshold flips may move a handful of episodes between classes
    assert abs(sc["episodes"] - ost.episodes) <= 8 and abs(sc["landed"] - ost.landed) <= 4


# =================================================================================================
# BASELINE configs[4]: curriculum sweep -- one episode per env per stage, device-reduced statistics
# =================================================================================================
def test_curriculum_sweep_matches_oracle():
    n = 3000
    caps = dd.step_schedule(4, 75, 250).tolist()
    env = dd.BatchedDroneEnv(n, device="cuda:0", seed=13, randomize_drone=True, randomize_platform=True,
                             auto_reset=False, dtype=torch.float64)
    stages = dd.curriculum_sweep(env, caps, policy="bangbang", reduce=False)
    o = co.OracleBatch(n, seed=13, randomize_drone=True, randomize_platform=True, auto_reset=False)
    for cap, st in zip(caps, stages):
        o.max_steps = cap
        o.reset()
        _, _, os_ = o.rollout(cap, policy=co.POL_BANGBANG)
        assert st["max_steps"] == cap and st["num_games"] == n == os_.episodes       # exactly one episode per env
        assert (st["num_successes"], st["crashed"], st["timed_out"]) == (os_.landed, os_.crashed, os_.truncated)
        assert st["avg_steps"] == os_.sum_length / n
        assert st["avg_reward"] == pytest.approx(os_.sum_return / n, rel=1e-9, abs=1e-5)
    rates = [s["success_rate"] for s in stages]
    assert rates == sorted(rates) and rates[-1] > rates[0]      # longer caps let more bang-bang episodes land
    with pytest.raises(ValueError):
        dd.collect_episodes(dd.BatchedDroneEnv(8, device="cuda:0", auto_reset=True), 10)


# =================================================================================================
# round 2: fp32 at size against the float64 oracle, and the device kernel against the HOST
# instantiation of its own source
# =================================================================================================
def test_f32_65536_envs_250_steps_vs_oracle():
    """BASELINE configs[3]'s population (65,536 envs x 250 steps, auto-reset off so that every env is compared for its
    whole episode) through the fp32 kernel against the float64 C oracle: north_star's 1e-5 on observations / rewards,
    flags and counters exact outside the BAND rule, flip rate reported and bounded."""
    n, T = 65536, 250
    i = np.arange(n)
    sx, sy = 100 + (i * 7919) % 601, 50 + (i * 104729) % 201
    spx, spy = 100 + (i * 1299709) % 600, 100 + (i * 15485863) % 450
    A = co.random_actions(21, 0, 0, T, n)
    A[:, 1::2] &= np.where(co.random_actions(22, 0, 0, T, n)[:, 1::2] < 2, 7, 1).astype(np.uint8)   # odd envs: side thrust rarely
    o = co.OracleBatch(n, randomize_drone=False, randomize_platform=False)
    o.inject(sx, sy, spx, spy)
    e = _env(n, dtype=torch.float32)
    e.inject(_t(sx, torch.float32), _t(sy, torch.float32), _t(spx, torch.float32), _t(spy, torch.float32))
    Ad = _t(A)
    valid = np.ones(n, bool)
    flips, worst = [], 0.0
    for t in range(T):
        oo, orr, od = o.step(A[t])
        obs, rew, fl = e.step_raw(Ad[t])
        fl = fl.cpu().numpy()
        bad = valid & (fl != od)
        if bad.any():
            margin = _threshold_margin(o)
            for j in np.nonzero(bad)[0]:
                assert margin[j] < BAND, f"env {j} step {t}: flags {fl[j]:#x} vs {od[j]:#x}, margin {margin[j]:.3g}"
                flips.append((int(j), t, float(margin[j])))
            valid &= ~bad
        if t % 10 == 9 or t == T - 1:                       # full compare every 10th step (the errors accumulate)
            assert np.array_equal(e.steps.cpu().numpy()[valid], o.steps[valid])
            obs, rew = obs.cpu().numpy(), rew.cpu().numpy()
            assert _close32(obs, oo)[valid].all(), f"step {t}: obs outside fp32 tolerance"
            assert _close32(rew, orr)[valid].all(), f"step {t}: reward outside fp32 tolerance"
            worst = max(worst, float(np.abs(obs - oo)[valid].max()))
    episodes = int((o.flags & co.DONE).astype(bool).sum())
    assert episodes > 0.8 * n
    assert len(flips) <= 8, flips                           # observed: 1-2 per run (~3e-5 per episode)
    print(f"fp32 at size: {len(flips)} flips in {episodes} episodes {flips}, worst normalised-obs abs error {worst:.3g}")


def test_device_kernel_vs_host_instantiation_of_the_same_source():
    """One step from identical states: step_kernel<float> on the GPU against csrc/drone_core.cuh instantiated for the
    host (tests/hosttwin.py).  Everything that does not pass through sin / cos must be BIT-IDENTICAL (same source, same
    explicit roundings); where sin / cos enter (main thruster fired, or the landing test ran) the two may differ by the
    device's sincospif rounding: <= 2 ulp of the velocity.  float64: identical up to libm-vs-CUDA sin / cos ulps."""
    from hosttwin import HostBatch
    n = 20000
    kw = dict(seed=31, randomize_drone=True, randomize_platform=True, max_steps=120, auto_reset=True)
    for dtype, npdt in ((torch.float32, np.float32), (torch.float64, np.float64)):
        e = dd.BatchedDroneEnv(n, device="cuda:0", dtype=dtype, **kw)
        e.reset()
        e.rollout(57, "random")                              # a population of mid-flight states
        h = HostBatch(n, npdt, **kw)
        for t in range(3):
            st = e.get_state()
            h.pos_vel[:] = e.pos_vel.cpu().numpy(); h.att_fuel[:] = e.att_fuel.cpu().numpy()
            h.platform[:] = e.platform.cpu().numpy(); h.steps[:] = e.steps.cpu().numpy()
            h.episode[:] = e.episode.cpu().numpy().astype(np.uint32); h.flags[:] = e.flags.cpu().numpy()
            act = e.random_actions(1, t0=1000 + t)[0]
            obs_d, rew_d, fl_d = e.step_raw(act)
            obs_h, rew_h, fl_h = h.step(act.cpu().numpy())
            obs_d, rew_d, fl_d = obs_d.cpu().numpy(), rew_d.cpu().numpy(), fl_d.cpu().numpy()
            a_np = act.cpu().numpy()
            no_trig = ((a_np & 1) == 0) & (fl_d == 0) & (fl_h == 0)        # main thruster off, still flying
            assert no_trig.sum() > n // 4
            if dtype == torch.float32:
                assert np.array_equal(obs_d[no_trig].view(np.uint32), obs_h[no_trig].view(np.uint32))
                assert np.array_equal(rew_d[no_trig].view(np.uint32), rew_h[no_trig].view(np.uint32))
                assert np.array_equal(e.pos_vel.cpu().numpy()[no_trig].view(np.uint32), h.pos_vel[no_trig].view(np.uint32))
                same = fl_d == fl_h
                assert (~same).sum() <= 2                                   # a threshold decision on a 1-ulp difference
                both = same & (fl_d == 0)
                assert np.abs(obs_d[both].astype(np.float64) - obs_h[both]).max() <= 3e-7    # ~2 ulp of v / 10
            else:
                assert np.array_equal(fl_d, fl_h)
                np.testing.assert_allclose(obs_d, obs_h, rtol=1e-14, atol=1e-15)
                np.testing.assert_allclose(rew_d, rew_h, rtol=1e-14, atol=1e-15)
