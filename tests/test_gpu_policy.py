"""K5 -- the fused policy rollout (tcgen05 MLP + env step in one kernel) on the GPU.

Tolerances.  The dense layers run on the tensor cores with bf16 operands (LayerNorm centring and gamma
folded into the weight images by dd_policy_pack; the first layer in split bf16, ~16 mantissa bits) and fp32
accumulation, so against the notebook's eager fp32 network the logits differ by up to ~2e-2 (measured 0.020
max on the fixture batch by an emulation of the same roundings in torch): probabilities are compared with
atol 1e-2, and thresholded actions may only differ where the fp32 |logit| < 0.05.  Against that emulation
(same roundings, different summation order) the kernel must agree to 3e-3 -- that is the check that the
UMMA descriptors / layouts / LayerNorm epilogues are right.  The critic (values of magnitude ~250) is
compared relative to that scale.
The environment half is exact: replaying the actions the fused kernel chose through dd_rollout
must reproduce rewards, flags and final state bit for bit.
"""
import importlib
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _eq(a, b):
    """torch.equal that treats NaN == NaN (prev_dist uses NaN for 'no previous state')."""
    if a.is_floating_point():
        a, b = torch.nan_to_num(a, nan=-12345.0), torch.nan_to_num(b, nan=-12345.0)
    return torch.equal(a, b)

from torch_reference import reference_policy

dd = importlib.import_module("reinforcement-learning-101_b200")
pol = importlib.import_module("reinforcement-learning-101_b200.policy")
nv = dd.native
DEV = "cuda:0"


@pytest.fixture(scope="module")
def fixture(golden_dir):
    d = np.load(os.path.join(golden_dir, "policy_v1.npz"))
    sd = {k: torch.from_numpy(d[k]) for k in d.files if k.startswith("network")}
    return d, sd


def _bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _fold(W, b, g):
    """dd_policy_pack's operand folding: gamma * (centred over the outputs) weights and bias, and 1/gamma."""
    gc = torch.where(g.abs() < 1e-12, torch.copysign(torch.full_like(g, 1e-12), g), g)
    Wc = (W.double() - W.double().mean(0, keepdim=True)).float()
    bc = (b.double() - b.double().mean()).float()
    return gc[:, None] * Wc, gc * bc, 1.0 / gc


def _hi_lo(v):
    hi = _bf(v)
    return hi + _bf(v - hi)


def _ln_folded(xpp, ig, be):
    n = xpp.shape[-1]
    var = ((xpp * ig) ** 2).sum(-1, keepdim=True) / n
    return xpp * torch.rsqrt(var + 1e-5) + be


def _split(t):
    hi = _bf(t)
    return hi, _bf(t - hi)


def _emulate_bf16(sd, x, head=3):
    """Same operand roundings as the kernel: LayerNorm centring and gamma folded into the weight images;
    first layer in split bf16 (x_hi W_hi + x_hi W_lo + x_lo W_hi, b0 riding as column 15 against a constant 1);
    b1/b2 as bf16 hi + lo; bf16 hidden activations and weights, fp32 accumulation, fp32 variance +
    normalisation, fp32 last layer.  head=3: sigmoid probabilities; head=1: the critic's raw value."""
    W0, b0, ig0 = _fold(sd["network.0.weight"], sd["network.0.bias"], sd["network.1.weight"])
    W1, b1, ig1 = _fold(sd["network.3.weight"], sd["network.3.bias"], sd["network.4.weight"])
    W2, b2, ig2 = _fold(sd["network.6.weight"], sd["network.6.bias"], sd["network.7.weight"])
    x16 = torch.cat([x, torch.ones(x.shape[0], 1)], 1)
    (xh, xl), (wh, wl) = _split(x16), _split(torch.cat([W0, b0[:, None]], 1))
    h = xh @ wh.T + xh @ wl.T + xl @ wh.T
    h = torch.relu(_ln_folded(h, ig0, sd["network.1.bias"]))
    h = _bf(h) @ _bf(W1).T + _hi_lo(b1)
    h = torch.relu(_ln_folded(h, ig1, sd["network.4.bias"]))
    h = _bf(h) @ _bf(W2).T + _hi_lo(b2)
    h = torch.relu(_ln_folded(h, ig2, sd["network.7.bias"]))
    z = h @ sd["network.9.weight"].T + sd["network.9.bias"]
    return torch.sigmoid(z) if head == 3 else z.squeeze(-1)


def test_forward_matches_torch_on_the_reference_checkpoint(fixture):
    d, sd = fixture
    blob = dd.PolicyBlob(sd, device=DEV)
    obs = torch.from_numpy(d["obs"]).to(DEV)
    probs = dd.policy_forward(blob, obs).cpu()
    emu = _emulate_bf16(sd, torch.from_numpy(d["obs"]))
    ref = torch.from_numpy(d["probs"])
    assert torch.isfinite(probs).all()
    assert (probs - emu).abs().max().item() < 3e-3, (probs - emu).abs().max().item()
    assert (probs - ref).abs().max().item() < 1e-2, (probs - ref).abs().max().item()
    flips = (probs > 0.5) != (ref > 0.5)
    assert np.abs(d["logits"])[flips.numpy()].max(initial=0.0) < 0.05
    # the module-based constructor gives the same blob
    blob2 = dd.PolicyBlob.from_module(reference_policy(sd), device=DEV)
    assert torch.equal(blob.blob, blob2.blob)


@pytest.mark.parametrize("n", [1, 127, 129, 512, 1000])
def test_forward_random_weights_ragged_sizes(n):
    g = torch.Generator().manual_seed(n)
    net = torch.nn.Sequential(
        torch.nn.Linear(15, 128), torch.nn.LayerNorm(128), torch.nn.ReLU(),
        torch.nn.Linear(128, 128), torch.nn.LayerNorm(128), torch.nn.ReLU(),
        torch.nn.Linear(128, 64), torch.nn.LayerNorm(64), torch.nn.ReLU(),
        torch.nn.Linear(64, 3), torch.nn.Sigmoid())
    with torch.no_grad():
        for p in net.parameters():                      # non-trivial LN affine and biases
            p.copy_(torch.randn(p.shape, generator=g) * (0.3 if p.dim() > 1 else 0.5) + (1.0 if p.dim() == 1 else 0.0) * 0.5)
    sd = {"network." + k: v for k, v in net.state_dict().items()}
    blob = dd.PolicyBlob(sd, device=DEV)
    x = torch.randn(n, 15, generator=g)
    probs = dd.policy_forward(blob, x.to(DEV)).cpu()
    emu = _emulate_bf16(sd, x)
    assert probs.shape == (n, 3)
    assert (probs - emu).abs().max().item() < 5e-3, (probs - emu).abs().max().item()
    with torch.no_grad():
        assert (probs - net(x)).abs().max().item() < 8e-2


def test_forward_degenerate_layernorm_gammas(fixture):
    """gamma is folded into the weight images and divided out again for the variance: zero, tiny, negative
    and large gammas must still give LayerNorm's result (a zero gamma makes that unit the constant beta)."""
    d, sd = fixture
    sd = {k_: v.clone() for k_, v in sd.items()}
    for key in ("network.1.weight", "network.4.weight", "network.7.weight"):
        g = sd[key]
        g[0] = 0.0; g[1] = -0.0; g[2] = 1e-20; g[3] = -3e-15; g[4] = -1.7; g[5] = 250.0; g[6] = 1e-6
    blob = dd.PolicyBlob(sd, device=DEV)
    x = torch.from_numpy(d["obs"])
    probs = dd.policy_forward(blob, x.to(DEV)).cpu()
    assert torch.isfinite(probs).all()
    with torch.no_grad():
        ref = reference_policy(sd)(x)
    assert (probs - _emulate_bf16(sd, x)).abs().max().item() < 5e-3
    assert (probs - ref).abs().max().item() < 6e-2


def test_rollout_env_half_is_exact_and_buffers_are_consistent(fixture):
    d, sd = fixture
    blob = dd.PolicyBlob(sd, device=DEV)
    n, T = 1000, 70
    kw = dict(seed=21, randomize_drone=True, randomize_platform=True, max_steps=50, auto_reset=True,
              dtype=torch.float32, env_id_base=5000)
    a = dd.BatchedDroneEnv(n, device=DEV, **kw); a.reset()
    b = dd.BatchedDroneEnv(n, device=DEV, **kw); b.reset()
    out = dd.policy_rollout(a, blob, T, sample=True, t0=3, want="arldop")
    # (1) replay the chosen actions through the plain rollout kernel: identical environment evolution
    rew = torch.empty(T, n, device=DEV); don = torch.empty(T, n, dtype=torch.uint8, device=DEV)
    obs = torch.empty(T, n, 15, device=DEV)
    b.rollout(T, "trace", actions=out["actions"], reward_out=rew, done_out=don, obs_out=obs)
    assert torch.equal(out["reward"], rew) and torch.equal(out["done"], don)
    for k_, v in a.get_state().items():
        assert _eq(v, b.get_state()[k_]), k_
    assert a.stats() == b.stats() and a.stats()["env_steps"] == n * T
    # (2) obs[t] is the observation the policy saw at step t: obs[0] = reset obs, obs[t+1] = post-step obs[t]
    c = dd.BatchedDroneEnv(n, device=DEV, **kw)
    assert torch.equal(out["obs"][0], c.reset())
    assert torch.equal(out["obs"][1:], obs[:-1])
    # (3) probs are what the stand-alone forward gives on those observations (same tensor-core path)
    for t in (0, 17, T - 1):
        assert torch.equal(out["probs"][t], dd.policy_forward(blob, out["obs"][t].contiguous()))
    # (4) log-prob = sum over thrusters of Bernoulli(probs).log_prob(action)  (c16:L61-63)
    bits = torch.stack([(out["actions"] >> j) & 1 for j in range(3)], -1).float()
    lp = torch.distributions.Bernoulli(probs=out["probs"]).log_prob(bits).sum(-1)
    assert torch.allclose(out["logp"], lp, rtol=1e-4, atol=2e-5)
    # (5) sampling follows the probabilities, is deterministic, and depends on t0
    p = out["probs"].double()
    freq, exp = bits.double().mean((0, 1)), p.mean((0, 1))
    sd_ = (p * (1 - p)).sum((0, 1)).sqrt() / (T * n)
    assert ((freq - exp).abs() < 6 * sd_).all(), (freq, exp)
    a2 = dd.BatchedDroneEnv(n, device=DEV, **kw); a2.reset()
    assert torch.equal(dd.policy_rollout(a2, blob, T, sample=True, t0=3, want="a")["actions"], out["actions"])
    a3 = dd.BatchedDroneEnv(n, device=DEV, **kw); a3.reset()
    assert not torch.equal(dd.policy_rollout(a3, blob, T, sample=True, t0=4, want="a")["actions"], out["actions"])
    # (6) the PPO-collection fast-path instantiation (exactly want="arldo", sampling, auto-reset, statistics: every switch
    # compile-time) computes what the generic one (taken above because of the probabilities output) computes
    a4 = dd.BatchedDroneEnv(n, device=DEV, **kw); a4.reset()
    fast = dd.policy_rollout(a4, blob, T, sample=True, t0=3, want="arldo")
    for key in ("actions", "reward", "logp", "done", "obs"):
        assert torch.equal(fast[key], out[key]), key
    assert a4.stats() == a.stats()
    for k_, v in a.get_state().items():
        assert _eq(v, a4.get_state()[k_]), k_


def test_threshold_rollout_lands_with_the_trained_policy(fixture):
    """The reference reports ~76 % landings for this checkpoint (README.md:78,108); random actions land ~3 %."""
    d, sd = fixture
    blob = dd.PolicyBlob(sd, device=DEV)
    n = 4096
    e = dd.BatchedDroneEnv(n, device=DEV, seed=1, randomize_drone=True, randomize_platform=True, max_steps=500,
                           auto_reset=True, dtype=torch.float32)
    e.reset()
    out = dd.policy_rollout(e, blob, 500, sample=False, want="ap")
    s = e.stats()
    assert s["episodes"] >= n and s["landing_rate"] > 0.3, s
    # thresholded actions are exactly probs > 0.5
    exp = sum(((out["probs"][..., j] > 0.5).to(torch.uint8) << j) for j in range(3))
    assert torch.equal(out["actions"], exp)
    print("trained policy through the fused kernel:", {k_: s[k_] for k_ in ("episodes", "landing_rate", "mean_return", "mean_length")})


def test_full_size_cfg4_properties_and_shard_invariance(fixture):
    """BASELINE configs[3] at its real size (65,536 envs x 250 steps): conservation laws of the statistics,
    replay of the chosen actions through dd_rollout (bit-identical rewards / flags / state), and independence of
    the launch shape -- two shards of N/2 with env_id_base 0 and N/2 give exactly the buffers of one env of N
    (Philox is keyed by the global env id; the MLP is row-independent)."""
    d, sd = fixture
    blob = dd.PolicyBlob(sd, device=DEV)
    n, T = 65536, 250
    kw = dict(seed=7, randomize_drone=True, randomize_platform=True, max_steps=250, auto_reset=True, dtype=torch.float32)
    a = dd.BatchedDroneEnv(n, device=DEV, **kw); a.reset()
    out = dd.policy_rollout(a, blob, T, sample=True, want="arld")
    s = a.stats()
    assert s["env_steps"] == n * T
    assert s["episodes"] == s["landed"] + s["crashed"] + s["truncated"]
    assert s["episodes"] == int(a.get_state()["episode"].sum().item()) - n
    assert int((out["done"] != 0).sum().item()) == s["episodes"]
    assert torch.isfinite(out["logp"]).all() and (out["logp"] <= 0).all()
    # replay through the plain T-step kernel
    b = dd.BatchedDroneEnv(n, device=DEV, **kw); b.reset()
    rew = torch.empty(T, n, device=DEV); don = torch.empty(T, n, dtype=torch.uint8, device=DEV)
    b.rollout(T, "trace", actions=out["actions"], reward_out=rew, done_out=don)
    assert torch.equal(out["reward"], rew) and torch.equal(out["done"], don)
    for k_, v in a.get_state().items():
        assert _eq(v, b.get_state()[k_]), k_
    assert a.stats() == b.stats()
    # two shards == one env
    h = n // 2
    parts = []
    for r in range(2):
        c = dd.BatchedDroneEnv(h, device=DEV, env_id_base=r * h, **kw); c.reset()
        parts.append((dd.policy_rollout(c, blob, T, sample=True, want="arld"), c))
    for key in ("actions", "reward", "logp", "done"):
        assert torch.equal(torch.cat([parts[0][0][key], parts[1][0][key]], dim=1), out[key]), key
    tot = {k_: parts[0][1].stats()[k_] + parts[1][1].stats()[k_] for k_ in ("episodes", "landed", "crashed", "truncated", "sum_length", "env_steps")}
    assert tot == {k_: s[k_] for k_ in tot}


@pytest.mark.parametrize("n", [1, 3, 130, 1001])
def test_rollout_ragged_sizes_use_the_fallback_obs_store(fixture, n):
    """n not a multiple of 4 (or a partial last tile) cannot use the TMA bulk store of the observation tile:
    the direct-store path must give the same observations as stepping the plain kernels."""
    d, sd = fixture
    blob = dd.PolicyBlob(sd, device=DEV)
    T = 40
    kw = dict(seed=3, randomize_drone=True, randomize_platform=True, max_steps=30, auto_reset=True, dtype=torch.float32)
    a = dd.BatchedDroneEnv(n, device=DEV, **kw); a.reset()
    out = dd.policy_rollout(a, blob, T, sample=True, want="ardo")
    b = dd.BatchedDroneEnv(n, device=DEV, **kw); first = b.reset().clone()
    obs = torch.empty(T, n, 15, device=DEV); rew = torch.empty(T, n, device=DEV)
    b.rollout(T, "trace", actions=out["actions"], reward_out=rew, obs_out=obs)
    assert torch.equal(out["obs"][0], first) and torch.equal(out["obs"][1:], obs[:-1]) and torch.equal(out["reward"], rew)


def test_critic_value_forward_and_rollout_values(fixture, golden_dir):
    """DroneTeacherBoi (same trunk, Linear(64, 1)) through the same tensor-core path: the reference's critic
    checkpoint against its eager fp32 values, the persistent forward over a [T, N, 15] rollout buffer (TMA-fed
    full tiles + ragged tail), and the [T+1, N] values that feed dd_gae."""
    d, sd = fixture
    c = np.load(os.path.join(golden_dir, "critic_v1.npz"))
    sdc = {k_: torch.from_numpy(c[k_]) for k_ in c.files if k_.startswith("network")}
    vblob = dd.ValueBlob(sdc, device=DEV)
    x = torch.from_numpy(c["obs"])
    ref = torch.from_numpy(c["values"])
    v = dd.value_forward(vblob, x.to(DEV)).cpu()
    emu = _emulate_bf16(sdc, x, head=1)
    scale = ref.std().item()
    assert v.shape == ref.shape and torch.isfinite(v).all()
    assert ((v - emu).abs() / (1 + emu.abs())).max().item() < 2e-3, ((v - emu).abs() / (1 + emu.abs())).max().item()
    assert (v - ref).abs().max().item() < 0.02 * scale and (v - ref).pow(2).mean().sqrt().item() < 0.003 * scale
    # persistent forward over a rollout buffer: same rows -> same values, whatever the launch shape
    blob = dd.PolicyBlob(sd, device=DEV)
    n, T = 1000, 37
    env = dd.BatchedDroneEnv(n, device=DEV, seed=5, randomize_drone=True, randomize_platform=True, max_steps=60,
                             auto_reset=True, dtype=torch.float32)
    env.reset()
    out = dd.policy_rollout(env, blob, T, sample=True, want="rdo")
    final_obs = env.observe().clone()
    vals = dd.rollout_values(vblob, out["obs"], final_obs)
    assert vals.shape == (T + 1, n)
    for t in (0, 11, T - 1):
        assert torch.equal(vals[t], dd.value_forward(vblob, out["obs"][t]))
    assert torch.equal(vals[T], dd.value_forward(vblob, final_obs))
    flat = dd.value_forward(vblob, out["obs"].view(-1, 15)[3:-5])            # unaligned start, ragged length
    assert torch.equal(flat, vals[:T].reshape(-1)[3:-5])
    with torch.no_grad():
        eager = reference_policy(sdc, head=1)(out["obs"].cpu().view(-1, 15)).squeeze(-1)
    assert (vals[:T].cpu().reshape(-1) - eager).abs().max().item() < 0.03 * scale
    # ... and straight into GAE + advantage normalisation (bit-exact against the notebook's loop: test_gpu_parity)
    adv = dd.normalize_advantages(dd.gae(out["reward"], vals, (out["done"] != 0).to(torch.uint8)), reduce=False)
    assert adv.shape == (T, n) and torch.isfinite(adv).all()
    # argument checks
    with pytest.raises(ValueError):
        dd.value_forward(blob, x.to(DEV))
    with pytest.raises(ValueError):
        dd.policy_forward(vblob, x.to(DEV))


def test_policy_forward_large_persistent(fixture):
    """More row blocks than SMs: the persistent forward (each CTA walks several blocks, next tile prefetched
    by TMA) must give the same probabilities as block-sized calls."""
    d, sd = fixture
    blob = dd.PolicyBlob(sd, device=DEV)
    g = torch.Generator().manual_seed(0)
    base = torch.from_numpy(d["obs"])
    x = (base[torch.randint(0, base.shape[0], (200_003,), generator=g)] * (1 + 0.01 * torch.randn(200_003, 15, generator=g))).to(DEV)
    big = dd.policy_forward(blob, x)
    for lo in (0, 65_536, 199_000):
        assert torch.equal(big[lo:lo + 1003], dd.policy_forward(blob, x[lo:lo + 1003].contiguous()))


def test_temperature_sampling_follows_the_notebook_formula(fixture):
    """evaluate_policy_simple (Actor_Critic_PPO.ipynb c18): adjusted = p**(1/t) / (p**(1/t) + (1-p)**(1/t)), actions
    ~ Bernoulli(adjusted); t = 0 is probs > 0.5.  The recorded probabilities stay the policy's own."""
    d, sd = fixture
    blob = dd.PolicyBlob(sd, device=DEV)
    n, T = 4096, 40
    kw = dict(seed=11, randomize_drone=True, randomize_platform=True, max_steps=100, auto_reset=True, dtype=torch.float32)
    for temp in (0.3, 2.5):
        e = dd.BatchedDroneEnv(n, device=DEV, **kw); e.reset()
        out = dd.policy_rollout(e, blob, T, sample=True, want="ap", temperature=temp)
        p = out["probs"].double()
        adj = p ** (1 / temp) / (p ** (1 / temp) + (1 - p) ** (1 / temp))
        bits = torch.stack([(out["actions"] >> j) & 1 for j in range(3)], -1).double()
        freq, exp = bits.mean((0, 1)), adj.mean((0, 1))
        sd_ = (adj * (1 - adj)).sum((0, 1)).sqrt() / (T * n)
        assert ((freq - exp).abs() < 6 * sd_).all(), (temp, freq, exp)
        assert ((freq - p.mean((0, 1))).abs() > 6 * sd_).any()         # ... and not the untempered distribution
    e = dd.BatchedDroneEnv(n, device=DEV, **kw); e.reset()
    out0 = dd.policy_rollout(e, blob, T, want="ap", temperature=0)
    assert torch.equal(out0["actions"], sum(((out0["probs"][..., j] > 0.5).to(torch.uint8) << j) for j in range(3)))
    with pytest.raises(ValueError):
        dd.policy_rollout(e, blob, T, temperature=-1.0)


def test_policy_argument_errors(fixture):
    d, sd = fixture
    blob = dd.PolicyBlob(sd, device=DEV)
    e64 = dd.BatchedDroneEnv(8, device=DEV, dtype=torch.float64); e64.reset()
    with pytest.raises(ValueError):
        dd.policy_rollout(e64, blob, 4)
    e = dd.BatchedDroneEnv(8, device=DEV)
    with pytest.raises(RuntimeError):
        dd.policy_rollout(e, blob, 4)
    with pytest.raises(KeyError):
        dd.PolicyBlob({"network.0.weight": torch.zeros(128, 15)}, device=DEV)
    bad = dict(sd); bad["network.3.weight"] = torch.zeros(64, 128)
    with pytest.raises(ValueError):
        dd.PolicyBlob(bad, device=DEV)
    L = nv.lib()
    assert L.dd_policy_forward(None, None, None, None, 4, None) == -1
    assert L.dd_policy_pack(None, None, None, None) == -1
