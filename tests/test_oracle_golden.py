"""Pins the oracle (oracle/drone_port.py and oracle/drone_oracle.c) against golden
vectors recorded from the UNMODIFIED reference DroneGame (tests/golden/make_golden.py).
CPU only."""
import json
import os

import numpy as np
import pytest

from corpus import N_CORPUS, N_TRAJ, T_CORPUS, corpus_actions, corpus_spawns
from oracle import c_oracle as co
from oracle.drone_port import PortDroneGame, obs_vector

POLICIES = {
    "KAT1_no_thrust": lambda g: 0,
    "KAT2_main": lambda g: 1,
    "KAT3_all": lambda g: 7,
    "KAT4_right": lambda g: 4,
    "KAT5_main_right": lambda g: 5,
    "KAT6_bangbang": lambda g: int(g["vy"] > 1.5),
}


def _bits_to_action(b):
    return {"main_thrust": b & 1, "left_thrust": (b >> 1) & 1, "right_thrust": (b >> 2) & 1}


@pytest.fixture(scope="module")
def kat(golden_dir):
    return json.load(open(os.path.join(golden_dir, "kat.json")))


@pytest.mark.parametrize("name", sorted(POLICIES))
def test_port_kat_bit_exact(kat, name):
    ref = kat[name]
    g = PortDroneGame(None, False, False)
    g.reset()
    total, r, state, info = 0.0, 0.0, None, None
    head = []
    while not g.done:
        state, r, _, info = g.step(_bits_to_action(POLICIES[name]({"vy": g.vy})))
        total += r
        if g.steps <= 3:
            head.append({"y": float(g.y), "r": float(r)})
    assert g.steps == ref["steps"]
    for k, v in (("last_reward", r), ("total_reward", total), ("x", g.x), ("y", g.y), ("vx", g.vx),
                 ("vy", g.vy), ("angle", g.angle), ("angvel", g.angvel), ("fuel", g.fuel)):
        assert float(v) == ref[k], k          # same numpy build => bit identical
    assert (g.landed, g.crashed) == (ref["landed"], ref["crashed"])
    assert head == ref["head"]
    for k, v in ref["final_state"].items():
        assert float(state[k]) == float(v), k
    for k, v in ref["final_info"].items():
        assert float(info[k]) == v, k
    s2, r2, d2, i2 = g.step({"main_thrust": 1})
    assert (r2, d2, s2["steps"], i2["needs_reset"]) == (0, True, ref["after_done"]["steps"], True)
    assert sorted(i2) == ref["after_done"]["info_keys"]


def test_port_seed42_spawn(kat):
    np.random.seed(42)
    g = PortDroneGame(None, True, True)
    g.reset()
    assert [g.x, g.y, g.px, g.py] == kat["KAT7_seed42_spawn"]


@pytest.mark.parametrize("name", sorted(POLICIES))
def test_c_oracle_kat(kat, name):
    ref = kat[name]
    b = co.OracleBatch(1, randomize_drone=False, randomize_platform=False)
    b.reset()
    total, r, flags = 0.0, None, 0
    for _ in range(3000):
        a = POLICIES[name]({"vy": b.vy[0]})
        obs, rew, done = b.step(np.array([a], np.uint8))
        total += rew[0]
        r, flags = rew[0], done[0]
        if flags & co.DONE:
            break
    assert b.steps[0] == ref["steps"]
    assert bool(flags & co.LANDED) == ref["landed"] and bool(flags & co.CRASHED) == ref["crashed"]
    tol = dict(rel=1e-13, abs=1e-12)       # libm vs numpy sin/cos: <= 1 ulp, amplified by accumulation
    for k, v in (("last_reward", r), ("total_reward", total), ("x", b.x[0]), ("y", b.y[0]), ("vx", b.vx[0]),
                 ("vy", b.vy[0]), ("angle", b.angle[0]), ("angvel", b.angvel[0]), ("fuel", b.fuel[0])):
        assert float(v) == pytest.approx(ref[k], **tol), k
    assert b.ep_return[0] == pytest.approx(ref["total_reward"], **tol)
    for j, k in enumerate(("drone_x", "drone_y", "drone_vx", "drone_vy", "drone_angle", "drone_angular_vel",
                           "drone_fuel", "platform_x", "platform_y", "distance_to_platform",
                           "dx_to_platform", "dy_to_platform", "speed", "landed", "crashed")):
        assert obs[0, j] == pytest.approx(float(ref["final_state"][k]), **tol), k
    # freeze after done
    obs2, rew2, done2 = b.step(np.array([1], np.uint8))
    assert rew2[0] == 0.0 and (done2[0] & co.DONE) and b.steps[0] == ref["steps"]
    assert np.array_equal(obs2, obs)


def test_c_oracle_corpus(golden_dir):
    """4096 envs x 250 steps against the reference: flags/steps exact, state to ~1e-12."""
    gs = np.load(os.path.join(golden_dir, "corpus_summary.npz"))
    gt = np.load(os.path.join(golden_dir, "corpus_traj.npz"))
    sx, sy, spx, spy = corpus_spawns()
    A = corpus_actions()
    b = co.OracleBatch(N_CORPUS, randomize_drone=False, randomize_platform=False)
    b.inject(sx, sy, spx, spy)
    done_step = np.zeros(N_CORPUS, np.int16)
    for t in range(T_CORPUS):
        obs, rew, done = b.step(A[t])
        d = (done & co.DONE) > 0
        np.testing.assert_allclose(obs.sum(0), gs["obs_sum"][t], rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(rew.sum(), gs["rew_sum"][t], rtol=1e-12, atol=1e-9)
        assert int(d.sum()) == gs["done_cnt"][t]
        np.testing.assert_allclose(obs[:N_TRAJ], gt["obs"][t], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(rew[:N_TRAJ], gt["reward"][t], rtol=1e-12, atol=1e-13)
        assert np.array_equal(d[:N_TRAJ], gt["done"][t].astype(bool))
        newly = d & (done_step == 0)
        done_step[newly] = b.steps[newly]
    assert np.array_equal(done_step, gs["done_step"])
    assert np.array_equal(b.flags & 7, gs["flags"])
    np.testing.assert_allclose(b.ep_return, gs["total"], rtol=1e-12, atol=1e-12)
    fin = np.stack([b.x, b.y, b.vx, b.vy, b.angle, b.angvel, b.fuel], 1)
    np.testing.assert_allclose(fin, gs["final"], rtol=1e-12, atol=1e-12)
    # the corpus exercises every termination cause
    causes = set((b.flags[(b.flags & co.DONE) > 0] & co.CAUSE_MASK).tolist())
    assert causes == {0, co.CAUSE_GROUND, co.CAUSE_FUEL, co.CAUSE_OOB} or causes >= {0, co.CAUSE_GROUND, co.CAUSE_OOB}


def test_port_matches_corpus_trajectories(golden_dir):
    """python port, bit-exact, on the stored per-step trajectories."""
    gt = np.load(os.path.join(golden_dir, "corpus_traj.npz"))
    sx, sy, spx, spy = corpus_spawns()
    A = corpus_actions()
    for i in range(N_TRAJ):
        g = PortDroneGame(None, False, False)
        g.reset()
        g.inject(int(sx[i]), int(sy[i]), int(spx[i]), int(spy[i]))
        for t in range(T_CORPUS):
            s, r, d, _ = g.step(_bits_to_action(int(A[t, i])))
            assert np.array_equal(obs_vector(s), gt["obs"][t, i]), (i, t)
            assert float(r) == gt["reward"][t, i] and bool(d) == bool(gt["done"][t, i])


PHILOX_KAT = [  # Random123 kat_vectors, philox4x32 10 rounds
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.mark.parametrize("ctr,key,out", PHILOX_KAT)
def test_philox_known_answers(ctr, key, out):
    assert co.philox(ctr, key) == out


def test_spawn_ranges_and_uniformity():
    """Philox spawns cover exactly the reference's integer ranges (game_engine.py:66-83)."""
    n = 200_000
    b = co.OracleBatch(n, seed=7, randomize_drone=True, randomize_platform=True)
    b.reset()
    for arr, lo, hi in ((b.x, 100, 700), (b.y, 50, 250), (b.px, 100, 699), (b.py, 100, 549)):
        assert arr.min() == lo and arr.max() == hi and np.all(arr == np.round(arr))
        cnt = np.bincount(arr.astype(np.int64) - lo, minlength=hi - lo + 1)
        exp = n / (hi - lo + 1)
        chi2 = ((cnt - exp) ** 2 / exp).sum()
        dof = hi - lo
        assert abs(chi2 - dof) < 6 * np.sqrt(2 * dof)
    assert np.all(b.episode == 1) and np.all(b.fuel == 1000.0)
    # fixed spawn when randomisation is off (config.py:61-62, 34)
    f = co.OracleBatch(4, randomize_drone=False, randomize_platform=False)
    f.reset()
    assert (f.x[0], f.y[0], f.px[0], f.py[0]) == (400.0, 100.0, 400.0, 500.0)


def test_fuel_gating_order():
    """drone.py:58-76: fuel is re-tested before each thruster; odd fuel clamps at 0."""
    b = co.OracleBatch(2, randomize_drone=False, randomize_platform=False)
    b.reset()
    b.fuel[:] = [2.0, 1.0]
    b.step(np.array([7, 7], np.uint8))
    # env0: main fires (fuel 2->0), left/right blocked -> angvel stays 0, then fuel<=0 crash
    assert b.fuel[0] == 0.0 and b.angvel[0] == 0.0 and (b.flags[0] & co.CAUSE_MASK) == co.CAUSE_FUEL
    # env1: main fires with fuel 1 (-> -1), sides blocked, clamp to 0
    assert b.fuel[1] == 0.0 and b.angvel[1] == 0.0 and b.vy[1] < 0.3 * 0.99


def test_auto_reset_and_truncation():
    b = co.OracleBatch(8, seed=3, randomize_drone=True, randomize_platform=True, max_steps=5, auto_reset=True)
    b.reset()
    ep0 = b.episode.copy()
    for t in range(5):
        obs, rew, done, fin = b.step(np.zeros(8, np.uint8), want_final=True)
    assert np.all(done & co.TRUNCATED) and np.all(done & co.DONE)
    assert np.all(b.steps == 0) and np.all(b.flags == 0) and np.all(b.episode == ep0 + 1)
    assert np.all(obs[:, 2:6] == 0) and np.all(obs[:, 6] == 1.0)      # fresh episode obs
    assert np.all(fin[:, 3] > 0)                                      # terminal obs still falling
    assert b.stats.episodes == 8 and b.stats.truncated == 8 and b.stats.sum_length == 40.0


def test_shaping_port_matches_the_notebook_cells(golden_dir):
    """oracle/shaping_port.py against outputs of the PPO notebook's own calc_reward cells
    (tests/golden/make_shaping_golden.py executes Actor_Critic_PPO.ipynb c6-c7): bit-exact."""
    from oracle.shaping_port import EpisodeShaper, shaped_reward
    d = np.load(os.path.join(golden_dir, "shaping_golden.npz"))
    n_checked = 0
    kinds = set()
    for e in range(d["obs"].shape[0]):
        sh = EpisodeShaper(int(d["max_steps"]))
        for k in range(int(d["n_steps"][e])):
            r, timed_out = sh.step(d["obs"][e, k], d["obs"][e, k + 1])
            assert r == d["reward"][e, k], (e, k)
            n_checked += 1
            nxt = d["obs"][e, k + 1]
            kinds.add("landed" if nxt[13] else "crashed" if nxt[14] else "timeout" if timed_out else "flying")
    assert n_checked > 3000 and kinds == {"landed", "crashed", "timeout", "flying"}
    # the policy-gradient notebook's stateless calc_reward(state) on the same episodes
    g = np.load(os.path.join(golden_dir, "shaping_pg_golden.npz"))
    n_pg = 0
    for e in range(d["obs"].shape[0]):
        sh = EpisodeShaper(int(d["max_steps"]), variant="pg")
        for k in range(int(d["n_steps"][e])):
            r, _ = sh.step(d["obs"][e, k], d["obs"][e, k + 1])
            assert r == g["reward"][e, k], (e, k, r, g["reward"][e, k])
            n_pg += 1
    assert n_pg == n_checked
