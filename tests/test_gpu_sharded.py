"""The product launch path (round 2): ``ShardedDroneEnv`` -- CUDA graphs over parallel chains of independent shards --
planned step launches (dd_step_plan / dd_step_planned), the per-call device guard of the C ABI, the env-owned noise
counter (t0 contract of include/drone_b200.h) and the prev_dist invalidation.  GPU only.

Everything here is an equality: the fast paths must be bit-identical to plain per-step dd_step calls."""
import ctypes as C
import importlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
dd = importlib.import_module("reinforcement-learning-101_b200")
nv = dd.native
DEV = "cuda:0"
KW = dict(seed=5, randomize_drone=True, randomize_platform=True, max_steps=40, auto_reset=True, dtype=torch.float32)


def _eq(a, b):
    if a.is_floating_point():
        a, b = torch.nan_to_num(a, nan=-12345.0), torch.nan_to_num(b, nan=-12345.0)
    return torch.equal(a, b)


def _same_state(ea, eb):
    sa, sb = ea.get_state(), eb.get_state()
    for k in sa:
        if k == "prev_dist":
            continue
        assert _eq(sa[k], sb[k]), k


@pytest.mark.parametrize("chains", [1, 2, 3])
def test_run_from_graphs_equals_eager_stepping(chains):
    """ShardedDroneEnv.run(k) for an awkward sequence of k (pieces cut at the schedule period, cached graphs replayed in
    a different order than captured) == the same launches issued one by one, == one BatchedDroneEnv over all S*N envs
    stepped with the same actions (Philox is keyed by the global env id)."""
    S, N, L = 3, 1000, 4
    g = dd.ShardedDroneEnv(S, N, device=DEV, chains=chains, trace_len=L, use_graphs=True, env_id_base=50, **KW)
    e = dd.ShardedDroneEnv(S, N, device=DEV, chains=1, trace_len=L, use_graphs=False, env_id_base=50, **KW)
    big = dd.BatchedDroneEnv(S * N, device=DEV, env_id_base=50, **KW)
    g.reset(); e.reset(); big.reset()
    tr = g.random_trace()
    e.set_trace(tr.clone())
    j = 0
    for k in (1, 5, 20, 20, 20, 7, 12, 1, 30, 20, 20):
        g.run(k); e.run(k)
        assert g.t == e.t == j + k
        j += k
    g.join(); e.join()
    torch.cuda.synchronize()
    assert g.graph_replays > 0 and g.eager_launches == 0 and e.graph_replays == 0
    for s in range(S):
        _same_state(g.shards[s], e.shards[s])
        for name in ("obs", "reward", "step_flags"):
            assert torch.equal(getattr(g.shards[s], name), getattr(e.shards[s], name)), (s, name)
    assert g.stats() == e.stats()
    # one big env: shard s advanced by (number of visits) steps with its trace rows in order
    visits = [len([q for q in range(j) if q % S == s]) for s in range(S)]
    assert len(set(visits)) == 1                                   # j is a multiple of S here
    for v in range(visits[0]):
        big.step_raw(tr[v % L].reshape(-1))
    st = big.get_state()
    cat = {k: torch.cat([g.shards[s].get_state()[k] for s in range(S)]) for k in st if k != "prev_dist"}
    for k in cat:
        assert _eq(cat[k], st[k]), k
    sb = big.stats()
    assert {k: sb[k] for k in ("episodes", "landed", "crashed", "truncated", "sum_length", "env_steps")} == \
           {k: g.stats()[k] for k in ("episodes", "landed", "crashed", "truncated", "sum_length", "env_steps")}
    assert sb["episodes"] > 0


def test_run_want_obs_false_and_max_steps_change_drop_graphs():
    S, N, L = 2, 513, 3
    g = dd.ShardedDroneEnv(S, N, device=DEV, chains=2, trace_len=L, **KW)
    e = dd.ShardedDroneEnv(S, N, device=DEV, chains=2, trace_len=L, use_graphs=False, **KW)
    g.reset(); e.reset()
    e.set_trace(g.random_trace().clone())
    g.run(9, want_obs=False); e.run(9, want_obs=False)
    g.max_steps = 7; e.max_steps = 7                                # curriculum knob: cached graphs embed the old cap
    assert g.graphs_cached == 0
    g.run(11); e.run(11)
    g.join(); e.join()
    for s in range(S):
        _same_state(g.shards[s], e.shards[s])
        assert torch.equal(g.shards[s].obs, e.shards[s].obs)
    assert int(max(sh.steps.max().item() for sh in g.shards)) <= 7


def test_step_all_policy_in_the_loop():
    """step_all: the caller rewrites the static action buffer each step; one graph replay advances every shard."""
    S, N = 4, 700
    g = dd.ShardedDroneEnv(S, N, device=DEV, chains=2, **KW)
    ref = [dd.BatchedDroneEnv(N, device=DEV, env_id_base=s * N, **KW) for s in range(S)]
    g.reset()
    for r in ref:
        r.reset()
    gen = torch.Generator(device=DEV).manual_seed(3)
    for t in range(25):
        a = torch.randint(0, 8, (S, N), dtype=torch.uint8, device=DEV, generator=gen)
        obs, rew, fl = g.step_all(a)
        for s in range(S):
            o2, r2, f2 = ref[s].step_raw(a[s])
            assert torch.equal(obs[s], o2) and torch.equal(rew[s], r2) and torch.equal(fl[s], f2), (t, s)
    assert g.graphs_cached == 1 and g.graph_replays == 25 * g.C


def test_planned_step_is_the_same_launch_as_dd_step():
    """Raw C ABI: dd_step_plan + dd_step_planned against dd_step, float32 and float64, with and without obs."""
    L = nv.lib()
    for dtype in (torch.float32, torch.float64):
        n = 777
        a = dd.BatchedDroneEnv(n, device=DEV, want_final_obs=True, **dict(KW, dtype=dtype))
        b = dd.BatchedDroneEnv(n, device=DEV, want_final_obs=True, **dict(KW, dtype=dtype))
        a.reset(); b.reset()
        st = torch.cuda.current_stream().cuda_stream
        acts = a.random_actions(30)
        for want_obs in (True, False):
            plan = nv.DDStepPlan()
            rc = L.dd_step_plan(C.byref(a._state), C.byref(a.params), C.byref(a._cfg), a.obs.data_ptr() if want_obs else None,
                                15, a.reward.data_ptr(), a.step_flags.data_ptr(), a.final_obs.data_ptr(), a.stats_slots.data_ptr(),
                                n, C.byref(plan))
            assert rc == 0
            for t in range(15):
                assert L.dd_step_planned(C.byref(plan), acts[t].data_ptr(), st) == 0
                assert L.dd_step(C.byref(b._state), C.byref(b.params), C.byref(b._cfg), acts[t].data_ptr(),
                                 b.obs.data_ptr() if want_obs else None, 15, b.reward.data_ptr(), b.step_flags.data_ptr(),
                                 b.final_obs.data_ptr(), b.stats_slots.data_ptr(), n, st) == 0
            assert torch.equal(a.reward, b.reward) and torch.equal(a.step_flags, b.step_flags) and torch.equal(a.obs, b.obs)
            assert torch.equal(a.final_obs, b.final_obs)
            _same_state(a, b)
        assert a.stats() == b.stats()
    # argument errors: unplanned storage, NULL actions, bad stride
    blank = nv.DDStepPlan()
    assert L.dd_step_planned(C.byref(blank), acts[0].data_ptr(), st) == -2
    assert L.dd_step_planned(C.byref(plan), None, st) == -1
    assert L.dd_step_plan(C.byref(a._state), C.byref(a.params), C.byref(a._cfg), a.obs.data_ptr(), 14, None, None, None, None, n,
                          C.byref(blank)) == -2
    assert L.dd_step_planned(C.byref(blank), acts[0].data_ptr(), st) == -2          # a failed plan stays unusable


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_env_on_a_non_current_device():
    """every native call must run on the device that owns the env, whatever the current device is."""
    torch.cuda.set_device(0)
    kw = dict(KW)
    a = dd.BatchedDroneEnv(3000, device="cuda:1", **kw)
    b = dd.BatchedDroneEnv(3000, device="cuda:0", **kw)
    a.reset(); b.reset()
    acts = b.random_actions(20)
    for t in range(20):
        oa, ra, fa = a.step_raw(acts[t].to("cuda:1"))
        ob, rb, fb = b.step_raw(acts[t])
        assert torch.equal(oa.cpu(), ob.cpu()) and torch.equal(fa.cpu(), fb.cpu())
    a.rollout(10, "random"); b.rollout(10, "random")
    assert a.stats() == b.stats()
    assert torch.cuda.current_device() == 0
    adv = dd.normalize_advantages(a.reward.clone(), reduce=False)
    assert adv.device == a.reward.device and torch.isfinite(adv).all()
    import os
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "policy_v1.npz"))
    sd = {k: torch.from_numpy(fx[k]) for k in fx.files if k.startswith("network")}
    b1, b0 = dd.PolicyBlob(sd, device="cuda:1"), dd.PolicyBlob(sd, device="cuda:0")
    o1 = dd.policy_rollout(a, b1, 8, want="ar")
    o0 = dd.policy_rollout(b, b0, 8, want="ar")
    assert torch.equal(o1["actions"].cpu(), o0["actions"].cpu()) and torch.equal(o1["reward"].cpu(), o0["reward"].cpu())
    p1 = dd.policy_forward(b1, torch.from_numpy(fx["obs"][:300]).to("cuda:1"))
    p0 = dd.policy_forward(b0, torch.from_numpy(fx["obs"][:300]).to("cuda:0"))
    assert torch.equal(p1.cpu(), p0.cpu())


def test_noise_counter_advances_between_rollouts(golden_dir):
    """ADVICE r1 / t0 contract: successive rollouts of one env draw fresh noise (the env owns the running offset);
    an explicit t0 reproduces a draw and leaves the counter alone."""
    import os
    n, T = 512, 20
    e = dd.BatchedDroneEnv(n, device=DEV, **KW); e.reset()
    a1 = torch.empty(T, n, dtype=torch.uint8, device=DEV)
    assert e.t_rollout == 0
    tr0 = e.random_actions(2 * T)                                # the stream rollout(policy='random') follows
    r1 = torch.empty(T, n, device=DEV); r2 = torch.empty(T, n, device=DEV)
    e.rollout(T, "random", reward_out=r1)
    assert e.t_rollout == T
    e.rollout(T, "random", reward_out=r2)
    assert e.t_rollout == 2 * T
    f = dd.BatchedDroneEnv(n, device=DEV, **KW); f.reset()
    q = torch.empty(2 * T, n, device=DEV)
    f.rollout(2 * T, "trace", actions=tr0, reward_out=q)         # one continuous stream == two consecutive rollouts
    assert torch.equal(torch.cat([r1, r2]), q)
    d = np.load(os.path.join(golden_dir, "policy_v1.npz"))
    blob = dd.PolicyBlob({k: torch.from_numpy(d[k]) for k in d.files if k.startswith("network")}, device=DEV)
    x = dd.BatchedDroneEnv(n, device=DEV, **KW); x.reset()
    o1 = dd.policy_rollout(x, blob, T, want="a")["actions"].clone()
    assert x.t_rollout == T
    y = dd.BatchedDroneEnv(n, device=DEV, **KW); y.reset()
    p1 = dd.policy_rollout(y, blob, T, want="a", t0=0)["actions"].clone()
    assert torch.equal(o1, p1) and y.t_rollout == 0
    # same env state, same t0 -> same actions; the running counter -> different noise
    z0 = x.get_state()
    o2 = dd.policy_rollout(x, blob, T, want="a")["actions"].clone()
    x.set_state(z0)
    o3 = dd.policy_rollout(x, blob, T, want="a", t0=T)["actions"].clone()
    x.set_state(z0)
    o4 = dd.policy_rollout(x, blob, T, want="a", t0=0)["actions"].clone()
    assert torch.equal(o2, o3) and not torch.equal(o2, o4)
    st = [dd.collect_episodes(dd.BatchedDroneEnv(64, device=DEV, **dict(KW, auto_reset=False)), 10, reduce=False)]
    assert st[0]["num_games"] == 64


def test_prev_dist_is_invalidated_by_plain_steps():
    """dd_step does not advance prev_dist, so a shaped rollout after plain steps must start from 'no
    previous state' (NaN), exactly like a fresh env put into the same state."""
    n = 300
    kw = dict(KW, max_steps=0, auto_reset=False)
    a = dd.BatchedDroneEnv(n, device=DEV, **kw); a.reset()
    sh = torch.empty(10, n, device=DEV)
    a.rollout(10, "random", shaped_out=sh)                        # prev_dist now holds real distances
    assert not a.prev_dist.isnan().all()
    acts = a.random_actions(5, t0=100)
    for t in range(5):
        a.step_raw(acts[t])
    b = dd.BatchedDroneEnv(n, device=DEV, **kw); b.reset()
    st = a.get_state(); st.pop("prev_dist")
    b.set_state(st)                                               # same state, prev_dist = NaN
    sa = torch.empty(6, n, device=DEV); sb = torch.empty(6, n, device=DEV)
    a.rollout(6, "bangbang", shaped_out=sa); b.rollout(6, "bangbang", shaped_out=sb)
    assert torch.equal(sa, sb)


@pytest.mark.parametrize("S,C,L,dtype", [(1, 2, 3, torch.float32), (2, 5, 1, torch.float64), (5, 2, 2, torch.float64)])
def test_sharded_edge_shapes(S, C, L, dtype):
    """One shard with two chains requested, more chains than shards, a one-row trace, the float64 instantiation: the graph
    path still equals eager stepping, and the captured pieces survive a set_trace() only for step_all."""
    N = 300
    kw = dict(KW, dtype=dtype)
    g = dd.ShardedDroneEnv(S, N, device=DEV, chains=C, trace_len=L, **kw)
    e = dd.ShardedDroneEnv(S, N, device=DEV, chains=C, trace_len=L, use_graphs=False, **kw)
    assert g.C == min(C, S)
    with pytest.raises(RuntimeError):
        g.run(1)                                              # no trace yet
    g.reset(); e.reset()
    e.set_trace(g.random_trace().clone())
    for k in (1, 2 * S * L + 1, 7):
        g.run(k); e.run(k)
    g.step_all(torch.zeros(S, N, dtype=torch.uint8, device=DEV)); e.step_all(torch.zeros(S, N, dtype=torch.uint8, device=DEV))
    n_all = sum(1 for key in g._graphs if key[0] == "all")
    g.set_trace(g.trace.clone())                               # drops the pieces captured over the old trace tensor only
    assert sum(1 for key in g._graphs if key[0] == "run") == 0 and sum(1 for key in g._graphs if key[0] == "all") == n_all == 1
    g.run(3); e.run(3)
    g.join(); e.join()
    for s in range(S):
        _same_state(g.shards[s], e.shards[s])
        assert torch.equal(g.shards[s].obs, e.shards[s].obs) and g.shards[s].obs.dtype == dtype
    assert g.stats() == e.stats()
    with pytest.raises(ValueError):
        g.set_trace(torch.zeros(L + 1, S, N, dtype=torch.uint8, device=DEV))
