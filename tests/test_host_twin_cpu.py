"""The kernels' OWN arithmetic source on the CPU box: csrc/drone_core.cuh (step_core, write_obs, spawn, shaped_reward,
Philox) instantiated for the host by csrc/host_twin.cpp (include/drone_b200_host.h) and checked against the golden
vectors recorded from the unmodified reference (/root/reference/delivery_drone/game/game_engine.py:95-216) and against
the float64 C oracle.  No GPU.

  * float64 twin: flags / counters bit-exact, continuous values to 1e-12 (libm vs numpy sin / cos ulps).
  * float32 twin: north_star's fp32 tolerance |a-b| <= 1e-5 max(|a|,1) on observations and rewards; flags bit-exact
    except envs whose float64 state is within BAND of a termination threshold (same rule as tests/test_gpu_parity.py).
"""
import os
import re

import numpy as np
import pytest

from corpus import N_CORPUS, N_TRAJ, T_CORPUS, corpus_actions, corpus_spawns
from hosttwin import HostBatch, lib, nv
from oracle import c_oracle as co
from oracle.shaping_port import EpisodeShaper

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BAND = 2e-4


def test_host_twin_exports_every_declared_symbol():
    h = open(os.path.join(ROOT, "include", "drone_b200_host.h")).read()
    declared = set(re.findall(r"^\s*int\s+(dd_\w+)\s*\(", h, re.M))
    assert declared == set(nv.HOST_TWIN_EXPORTS)
    L = lib()
    for sym in declared:
        assert hasattr(L, sym), sym
    assert L.dd_host_abi_version() == nv.ABI_VERSION


def _margin(o):
    rad = np.radians(o.angle)
    bx, by = o.x - 10.0 * np.sin(rad), o.y + 10.0 * np.cos(rad)
    speed = np.sqrt(o.vx ** 2 + o.vy ** 2)
    m = [np.abs(speed - 3.0), np.abs(np.abs(o.angle) - 20.0),
         np.abs(bx - (o.px - 50)), np.abs(bx - (o.px + 50)), np.abs(by - (o.py - 10)), np.abs(by - (o.py + 10)),
         np.abs(o.y - 550.0), np.abs(o.x + 50.0), np.abs(o.x - 850.0), np.abs(o.y + 50.0)]
    return np.min(np.stack(m), axis=0)


def test_f64_twin_corpus_vs_reference_golden(golden_dir):
    gs = np.load(os.path.join(golden_dir, "corpus_summary.npz"))
    gt = np.load(os.path.join(golden_dir, "corpus_traj.npz"))
    sx, sy, spx, spy = corpus_spawns()
    A = corpus_actions()
    h = HostBatch(N_CORPUS, np.float64, randomize_drone=False, randomize_platform=False)
    h.inject(sx, sy, spx, spy)
    o = co.OracleBatch(N_CORPUS, randomize_drone=False, randomize_platform=False)
    o.inject(sx, sy, spx, spy)
    done_step = np.zeros(N_CORPUS, np.int16)
    for t in range(T_CORPUS):
        obs, rew, fl = h.step(A[t])
        oo, orr, od = o.step(A[t])
        assert np.array_equal(fl, od), f"flags differ from the oracle at step {t}"
        assert np.array_equal(h.steps, o.steps)
        np.testing.assert_allclose(obs, oo, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(rew, orr, rtol=1e-12, atol=1e-12)
        # golden vectors recorded from the unmodified reference
        np.testing.assert_allclose(obs.sum(0), gs["obs_sum"][t], rtol=1e-11, atol=1e-8)
        np.testing.assert_allclose(obs[:N_TRAJ], gt["obs"][t], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(rew[:N_TRAJ], gt["reward"][t], rtol=1e-12, atol=1e-12)
        d = (fl & nv.DONE) > 0
        assert int(d.sum()) == gs["done_cnt"][t]
        newly = d & (done_step == 0)
        done_step[newly] = h.steps[newly]
    assert np.array_equal(done_step, gs["done_step"])
    assert np.array_equal(h.flags & 7, gs["flags"])
    np.testing.assert_allclose(h.att_fuel[:, 3], gs["total"], rtol=1e-12, atol=1e-11)
    fin = np.concatenate([h.pos_vel, h.att_fuel[:, :3]], 1)
    np.testing.assert_allclose(fin, gs["final"], rtol=1e-12, atol=1e-11)


def test_f32_twin_corpus_vs_oracle():
    sx, sy, spx, spy = corpus_spawns()
    A = corpus_actions()
    h = HostBatch(N_CORPUS, np.float32, randomize_drone=False, randomize_platform=False)
    h.inject(sx, sy, spx, spy)
    o = co.OracleBatch(N_CORPUS, randomize_drone=False, randomize_platform=False)
    o.inject(sx, sy, spx, spy)
    valid = np.ones(N_CORPUS, bool)
    flips, worst = [], 0.0
    for t in range(T_CORPUS):
        obs, rew, fl = h.step(A[t])
        oo, orr, od = o.step(A[t])
        bad = valid & (fl != od)
        if bad.any():
            mg = _margin(o)
            for i in np.nonzero(bad)[0]:
                assert mg[i] < BAND, f"env {i} step {t}: flags {fl[i]:#x} vs {od[i]:#x}, float64 margin {mg[i]:.3g}"
                flips.append((int(i), t, float(mg[i])))
            valid &= ~bad
        assert np.array_equal(h.steps[valid], o.steps[valid])
        err = np.abs(obs.astype(np.float64) - oo)
        assert (err <= 1e-5 * np.maximum(np.abs(oo), 1.0))[valid].all(), f"step {t}: obs outside fp32 tolerance"
        assert (np.abs(rew - orr) <= 1e-5 * np.maximum(np.abs(orr), 1.0))[valid].all()
        worst = max(worst, float(err[valid].max()))
    assert len(flips) <= 2, flips
    assert worst < 5e-6


@pytest.mark.parametrize("n", [1, 257, 3000])
def test_f64_twin_autoreset_philox_truncation_vs_oracle(n):
    kw = dict(seed=11, randomize_drone=True, randomize_platform=True, max_steps=60, auto_reset=True, env_id_base=1000)
    o = co.OracleBatch(n, **kw)
    o.reset()
    h = HostBatch(n, np.float64, **kw)
    obs0 = h.reset()
    assert np.array_equal(h.pos_vel[:, 0], o.x) and np.array_equal(h.platform[:, 1], o.py)      # Philox spawns
    np.testing.assert_allclose(obs0, o.write_obs_into(np.zeros((n, 15))), rtol=1e-15)
    T = 150
    A = co.random_actions(11, 1000, 0, T, n)
    for t in range(T):
        oo, orr, od, of = o.step(A[t], want_final=True)
        obs, rew, fl, fin = h.step(A[t], want_final=True)
        assert np.array_equal(fl, od), t
        np.testing.assert_allclose(obs, oo, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(rew, orr, rtol=1e-12, atol=1e-12)
        d = (fl & nv.DONE) > 0
        np.testing.assert_allclose(fin[d], of[d], rtol=1e-12, atol=1e-12)
        assert np.array_equal(h.steps, o.steps) and np.array_equal(h.episode, o.episode)
    s = h.stats
    assert (int(s[0]), int(s[1]), int(s[2]), int(s[3])) == (o.stats.episodes, o.stats.landed, o.stats.crashed, o.stats.truncated)
    assert int(s[5]) == o.stats.sum_length and int(s[0]) > 0
    assert np.int64(s[4]) / nv.RETURN_FIXED_SCALE == pytest.approx(o.stats.sum_return, abs=1e-6 * int(s[0]))    # 2^-20 fixed point per episode
    # the T-steps twin (in-kernel Philox action source, same orchestration as rollout_kernel) == stepping
    h2 = HostBatch(n, np.float64, **kw)
    h2.reset()
    out = h2.rollout(T, nv.POLICY_RANDOM, t0=0)
    for a_, b_ in ((h.pos_vel, h2.pos_vel), (h.att_fuel, h2.att_fuel), (h.platform, h2.platform), (h.steps, h2.steps),
                   (h.episode, h2.episode), (h.flags, h2.flags)):
        assert np.array_equal(a_, b_)
    assert np.array_equal(h.stats, h2.stats) and out["done"].any()


@pytest.mark.parametrize("variant", ["ppo", "pg"])
def test_f64_twin_shaped_reward_vs_notebook_port(variant):
    """shaped_reward_ppo / shaped_reward_pg of drone_core.cuh (the fused N2 epilogue) against oracle/shaping_port.py, which
    is pinned bit for bit to the notebooks' executed cells; freeze-after-done so that every terminal observation is in
    the stream."""
    n, T, ms = 64, 130, 100
    h = HostBatch(n, np.float64, seed=4, randomize_drone=True, randomize_platform=True, max_steps=ms, auto_reset=False,
                  shaping=variant)
    obs0 = h.reset()
    out = h.rollout(T, nv.POLICY_BANGBANG, want=("reward", "done", "obs", "shaped"))
    obs, don, got = out["obs"], out["done"], out["shaped"]
    exp = np.zeros((T, n))
    for i in range(n):
        sh = EpisodeShaper(ms, variant=variant)
        cur = obs0[i]
        for t in range(T):
            exp[t, i], _ = sh.step(cur, obs[t, i])
            cur = obs[t, i]
            if don[t, i]:
                break
    np.testing.assert_allclose(got, exp, rtol=1e-9, atol=1e-9)
    assert (don & nv.LANDED).any() and (don & nv.TRUNCATED).any()
    assert got.max() > (800 if variant == "ppo" else 500)


def test_twin_argument_errors():
    h = HostBatch(4, np.float64)
    L = lib()
    import ctypes as C
    assert L.dd_step_host(None, C.byref(h.params), C.byref(h.cfg), None, None, 15, None, None, None, None, 4) == -1
    assert L.dd_step_host(C.byref(h.state), C.byref(h.params), C.byref(h.cfg), None, None, 15, None, None, None, None, 4) == -1
    a = np.zeros(4, np.uint8); obs = np.zeros((4, 15))
    assert L.dd_step_host(C.byref(h.state), C.byref(h.params), C.byref(h.cfg), a.ctypes.data, obs.ctypes.data, 14, None, None, None, None, 4) == -2
    assert L.dd_step_host(C.byref(h.state), C.byref(h.params), C.byref(h.cfg), a.ctypes.data, None, 15, None, None, None, None, 0) == 0
