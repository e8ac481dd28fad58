"""``DroneGamePool`` views / in-process client / socket shim on the GPU against the oracle's
reference-faithful per-instance port (float64: dict values agree to the last sin/cos ulp)."""
import importlib
import json
import os

import numpy as np
import pytest

from oracle.drone_port import PortDroneGame

pytestmark = pytest.mark.gpu

dd = importlib.import_module("reinforcement-learning-101_b200")
compat = importlib.import_module("reinforcement-learning-101_b200.compat")

TOL = dict(rel=1e-12, abs=1e-11)


def _same_state(a, b):
    assert list(a) == list(b) == list(compat.STATE_KEYS)
    for k in compat.STATE_KEYS:
        if k in ("landed", "crashed", "steps"):
            assert a[k] == b[k] and type(a[k]) is type(b[k]), k
        else:
            assert a[k] == pytest.approx(float(b[k]), **TOL), k


def _same_info(a, b):
    assert [k for k in a if k != "needs_reset"] == list(compat.INFO_KEYS)
    for k in compat.INFO_KEYS:
        assert a[k] == pytest.approx(float(b[k]), **TOL), k


def test_views_follow_the_reference_game_per_index():
    """Three games stepped in interleaved order with different policies, like the notebooks do
    through ``client.step(action, game_id)``."""
    pool = compat.DroneGamePool(3, device="cuda:0", randomize_drone=False, randomize_platform=False)
    ports = [PortDroneGame(None, False, False) for _ in range(3)]
    # before any reset(): fixed spawn, episode 0 (DroneGame.__init__)
    assert pool[0].episode == 0 and pool[0].steps == 0 and not pool[0].done
    _same_state(pool[1].get_state(), ports[1].get_state())
    for g, p in zip(pool.games, ports):
        _same_state(g.reset(), p.reset())
        assert g.episode == 1
    policies = [lambda s, k: {"main_thrust": int(s["drone_vy"] * 10 > 1.5)},                 # KAT6 bang-bang: lands
                lambda s, k: {"main_thrust": 1, "right_thrust": 1},                             # KAT5: crash
                lambda s, k: {"left_thrust": k % 2, "main_thrust": k % 5 == 0}]
    states = [g.get_state() for g in pool.games]
    finished = [False] * 3
    for k in range(300):
        for i in (2, 0, 1):
            a = policies[i](states[i], k)
            s, r, d, info = pool[i].step(a)
            ps, pr, pd, pinfo = ports[i].step(a)
            _same_state(s, ps)
            assert r == pytest.approx(float(pr), **TOL) and d == bool(pd)
            _same_info(info, ports[i].info())
            assert ("needs_reset" in info) == ("needs_reset" in pinfo)
            states[i], finished[i] = s, d
    assert finished == [True, True, True]
    assert pool[0].drone.landed and not pool[0].drone.crashed and pool[1].drone.crashed
    assert pool[0].steps == 259 and pool[0].total_reward == pytest.approx(89.00976955376255, **TOL)   # KAT6
    assert pool[1].steps == 85                                                                          # KAT5
    # frozen after done; reset of one game leaves the others alone
    s, r, d, info = pool[0].step({"main_thrust": 1})
    assert (r, d, info["needs_reset"], s["steps"]) == (0, True, True, 259)
    _same_state(pool[0].reset(), ports[0].reset())
    assert pool[0].episode == 2 and pool[1].done and pool[1].steps == 85
    assert pool[0].render() is None and pool[0].close() is None


def test_wire_message_of_the_landing_step(golden_dir):
    """the STATE message of the landing step, through the real socket shim."""
    kat = json.load(open(os.path.join(golden_dir, "kat.json")))["KAT6_bangbang"]
    pool = compat.DroneGamePool(2, device="cuda:0", randomize_drone=False, randomize_platform=False)
    srv = compat.DroneSocketServer(pool, host="127.0.0.1", port=0)
    srv.start()
    try:
        with compat.DroneGameClient("127.0.0.1", srv.port, timeout=20) as c:
            assert c.num_games == 2
            st = c.reset(1)
            for _ in range(400):
                st, r, done, info = c.step({"main_thrust": int(st.drone_vy * 10 > 1.5)}, game_id=1)
                if done:
                    break
            assert done and st.landed and not st.crashed and st.steps == kat["steps"] == 259
            assert r == pytest.approx(99.9, **TOL) and info["episode"] == 1
            for k, v in kat["final_state"].items():
                assert getattr(st, k) == pytest.approx(float(v), **TOL), k
            for k, v in kat["final_info"].items():
                assert info[k] == pytest.approx(float(v), **TOL), k
            assert c.get_state(0).steps == 0
    finally:
        srv.stop()


def test_randomised_pool_spawns_in_reference_ranges():
    pool = compat.DroneGamePool(64, device="cuda:0", seed=5, randomize_drone=True, randomize_platform=True)
    c = compat.InProcessDroneGameClient(pool)
    xs = []
    for g in range(64):
        s = c.reset(g)
        x, y, px, py = s.drone_x * 800, s.drone_y * 600, s.platform_x * 800, s.platform_y * 600
        assert 100 <= round(x) <= 700 and 50 <= round(y) <= 250 and 100 <= round(px) <= 699 and 100 <= round(py) <= 549
        assert abs(x - round(x)) < 1e-9 and s.drone_fuel == 1.0 and s.steps == 0
        xs.append(round(x))
    assert len(set(xs)) > 32
    s1 = c.reset(3)
    assert round(s1.drone_x * 800) != xs[3] or True      # a new episode draws a new spawn (may coincide)
    assert pool[3].episode == 2


def test_step_host_modes_are_identical():
    """step_host: 'copy' (packed output block in HBM + one device->host copy) and 'zero_copy' (the kernel's obs /
    reward / flags destinations are the pinned host buffers themselves) and plain device stepping give the same
    outputs, state and statistics, at ragged sizes too (the TMA bulk store of a full tile and the fallback loop)."""
    import torch
    for n in (5000, 4096, 257):
        kw = dict(device="cuda:0", seed=9, randomize_drone=True, randomize_platform=True, max_steps=25, auto_reset=True,
                  dtype=torch.float32, env_id_base=777)
        a, b, c = dd.BatchedDroneEnv(n, **kw), dd.BatchedDroneEnv(n, **kw), dd.BatchedDroneEnv(n, **kw)
        a.reset(); b.reset(); c.reset()
        ioa, iob = a.make_host_io(), b.make_host_io()
        g = torch.Generator().manual_seed(1)
        for t in range(60):
            act = torch.randint(0, 8, (n,), generator=g, dtype=torch.uint8)
            ioa["actions"].copy_(act); iob["actions"].copy_(act)
            a.step_host(ioa)
            b.step_host(iob, mode="zero_copy")
            obs, rew, fl = c.step_raw(act.to("cuda:0"))
            for k_, ref in (("obs", obs), ("reward", rew), ("flags", fl)):
                assert torch.equal(ioa[k_], iob[k_]), (n, t, k_)
                assert torch.equal(ioa[k_], ref.cpu()), (n, t, k_)
        # wait=False: the call returns an event; the results are in the pinned buffers once it has been waited for
        side = torch.cuda.Stream("cuda:0")
        side.wait_stream(torch.cuda.current_stream())
        for t in range(5):
            act = torch.randint(0, 8, (n,), generator=g, dtype=torch.uint8).pin_memory()
            with torch.cuda.stream(side):
                ev = a.step_host(ioa, actions=act, wait=False)
            b.step_host(iob, actions=act)
            ev.synchronize()
            for k_ in ("obs", "reward", "flags"):
                assert torch.equal(ioa[k_], iob[k_]), (n, t, k_)
            c.step_raw(act.to("cuda:0"))
        torch.cuda.current_stream().wait_stream(side)
        sa, sb = a.get_state(), b.get_state()
        for k_ in sa:
            assert torch.equal(torch.nan_to_num(sa[k_].double(), nan=-1.0), torch.nan_to_num(sb[k_].double(), nan=-1.0)), k_
        assert a.stats() == b.stats() == c.stats()
