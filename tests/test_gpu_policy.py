"""K5 -- the fused policy rollout (tcgen05 MLP + env step in one kernel) on the GPU.

Tolerances.  The dense layers run on the tensor cores with 16-bit operands (LayerNorm centring, sign(gamma) and
the previous layer's |gamma| folded into the weight images by dd_policy_pack; the first layer in split 16-bit) and
fp32 accumulation.  The operand format is fp16 when the network provably fits its range (the reference checkpoints
do), bf16 otherwise.  Against the notebook's eager fp32 network: fp16 -- probabilities within 2e-3 (fixture) / 5e-3
(all 16.4 M rows of a cfg-4 rollout), thresholded actions differ only where |logit| < 0.02; bf16 -- 1e-2 / |logit| <
0.05 (round 1's only mode).  Against an emulation of the same roundings in torch (different summation order) the
kernel must agree to 5e-4 (fp16) / 3e-3 (bf16) -- that is the check that the UMMA descriptors / layouts / LayerNorm
epilogues are right.  The critic (values of magnitude ~250) is compared in absolute value units.
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


def _q(t, f16):
    """Round to the 16-bit operand format of the blob and back."""
    return t.to(torch.float16 if f16 else torch.bfloat16).to(torch.float32)


def _gabs(g):
    return g.abs().clamp_min(1e-12)


def _fold(W, b, g, gprev):
    """dd_policy_pack's operand folding: s (W - column mean over the outputs) g_prev and s (b - mean b), with
    s = sign(gamma) of THIS layer's LayerNorm and g_prev = |gamma| of the PREVIOUS one (None: the first layer)."""
    s = torch.where(g < 0, -torch.ones_like(g), torch.ones_like(g))
    Wc = W - W.double().mean(0, keepdim=True).float()
    bc = b - b.double().mean().float()
    img = s[:, None] * Wc
    if gprev is not None:
        img = img * _gabs(gprev)[None, :]
    return img, s * bc


def _split(t, f16):
    hi = _q(t, f16)
    return hi, _q(t - hi, f16)


def _ln_folded(xp, beta_over_g):
    n = xp.shape[-1]
    var = (xp ** 2).sum(-1, keepdim=True) / n
    return xp * torch.rsqrt(var + 1e-5) + beta_over_g


def _emulate16(sd, x, head=3, f16=True):
    """Same operand roundings as the kernel: LayerNorm centring, the sign of gamma and the previous layer's |gamma|
    folded into the weight images; first layer in split 16-bit (x_hi W_hi + x_hi W_lo + x_lo W_hi, b0 riding as column
    15 against a constant 1); b1/b2 as hi + lo; 16-bit hidden activations and weights (fp16 or bf16, as the blob says),
    fp32 accumulation, fp32 variance + normalisation, fp32 last layer (x |gamma| of the last LayerNorm).
    head=3: sigmoid probabilities; head=1: the critic's raw value."""
    g0, g1, g2 = sd["network.1.weight"], sd["network.4.weight"], sd["network.7.weight"]
    W0, b0 = _fold(sd["network.0.weight"], sd["network.0.bias"], g0, None)
    W1, b1 = _fold(sd["network.3.weight"], sd["network.3.bias"], g1, g0)
    W2, b2 = _fold(sd["network.6.weight"], sd["network.6.bias"], g2, g1)
    x16 = torch.cat([x, torch.ones(x.shape[0], 1)], 1)
    (xh, xl), (wh, wl) = _split(x16, f16), _split(torch.cat([W0, b0[:, None]], 1), f16)
    h = xh @ wh.T + xh @ wl.T + xl @ wh.T
    h = torch.relu(_ln_folded(h, sd["network.1.bias"] / _gabs(g0)))
    h = _q(h, f16) @ _q(W1, f16).T + sum(_split(b1, f16))
    h = torch.relu(_ln_folded(h, sd["network.4.bias"] / _gabs(g1)))
    h = _q(h, f16) @ _q(W2, f16).T + sum(_split(b2, f16))
    h = torch.relu(_ln_folded(h, sd["network.7.bias"] / _gabs(g2)))
    z = h @ (sd["network.9.weight"] * _gabs(g2)[None, :]).T + sd["network.9.bias"]
    return torch.sigmoid(z) if head == 3 else z.squeeze(-1)


def test_forward_matches_torch_on_the_reference_checkpoint(fixture):
    d, sd = fixture
    blob = dd.PolicyBlob(sd, device=DEV)
    obs = torch.from_numpy(d["obs"]).to(DEV)
    probs = dd.policy_forward(blob, obs).cpu()
    assert blob.operand_dtype == "fp16"                 # the reference checkpoint fits the fp16 range
    emu = _emulate16(sd, torch.from_numpy(d["obs"]), f16=True)
    ref = torch.from_numpy(d["probs"])
    assert torch.isfinite(probs).all()
    assert (probs - emu).abs().max().item() < 5e-4, (probs - emu).abs().max().item()
    assert (probs - ref).abs().max().item() < 2e-3, (probs - ref).abs().max().item()
    flips = (probs > 0.5) != (ref > 0.5)
    assert np.abs(d["logits"])[flips.numpy()].max(initial=0.0) < 0.01
    # the bf16 images (what a network outside the fp16 range gets) through the same kernels
    blob_b = dd.PolicyBlob(sd, device=DEV, operands="bf16")
    assert blob_b.operand_dtype == "bf16"
    probs_b = dd.policy_forward(blob_b, obs).cpu()
    assert (probs_b - _emulate16(sd, torch.from_numpy(d["obs"]), f16=False)).abs().max().item() < 3e-3
    assert (probs_b - ref).abs().max().item() < 1e-2
    flips = (probs_b > 0.5) != (ref > 0.5)
    assert np.abs(d["logits"])[flips.numpy()].max(initial=0.0) < 0.05
    print(f"policy fixture: max prob error vs eager fp32: fp16 operands {(probs - ref).abs().max().item():.2e}, "
          f"bf16 operands {(probs_b - ref).abs().max().item():.2e}")
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
    f16 = blob.operand_dtype == "fp16"
    emu = _emulate16(sd, x, f16=f16)
    assert probs.shape == (n, 3)
    assert (probs - emu).abs().max().item() < (1e-3 if f16 else 5e-3), (probs - emu).abs().max().item()
    with torch.no_grad():
        assert (probs - net(x)).abs().max().item() < (1e-2 if f16 else 8e-2)
    pb = dd.policy_forward(dd.PolicyBlob(sd, device=DEV, operands="bf16"), x.to(DEV)).cpu()
    assert (pb - _emulate16(sd, x, f16=False)).abs().max().item() < 5e-3


def test_forward_degenerate_layernorm_gammas(fixture):
    """gamma is folded into the weight images and divided out again for the variance: zero, tiny, negative
    and large gammas must still give LayerNorm's result (a zero gamma makes that unit the constant beta)."""
    d, sd = fixture
    sd = {k_: v.clone() for k_, v in sd.items()}
    for key in ("network.1.weight", "network.4.weight", "network.7.weight"):
        g = sd[key]
        g[0] = 0.0; g[1] = -0.0; g[2] = 1e-20; g[3] = -3e-15; g[4] = -1.7; g[5] = 250.0; g[6] = 1e-6
    blob = dd.PolicyBlob(sd, device=DEV)
    assert blob.operand_dtype == "bf16"                  # |gamma| < 2^-6: outside the fp16 range -> bf16 images
    with pytest.raises(nv.NativeError):
        dd.PolicyBlob(sd, device=DEV, operands="fp16")   # demanding fp16 for it is refused
    x = torch.from_numpy(d["obs"])
    probs = dd.policy_forward(blob, x.to(DEV)).cpu()
    assert torch.isfinite(probs).all()
    with torch.no_grad():
        ref = reference_policy(sd)(x)
    assert (probs - _emulate16(sd, x, f16=False)).abs().max().item() < 5e-3
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


@pytest.mark.parametrize("n", [148 * 3 * 128, 148 * 3 * 128 + 128, 60000, 148 * 4 * 128 + 4])
def test_launch_shapes_give_the_same_rollout(fixture, n):
    """The launcher spreads the tiles over the SMs up to 3 per SM and packs four to a CTA above that
    (dd_policy_rollout_grid); more than 4 tiles per SM run as waves.  Around those thresholds (and with a ragged last
    tile that leaves by TMA) a rollout of n envs must contain, bit for bit, the rollout of any sub-range of its env ids
    launched on its own (a small batch: one tile per CTA), observations included."""
    d, sd = fixture
    blob = dd.PolicyBlob(sd, device=DEV)
    T = 24
    kw = dict(seed=5, randomize_drone=True, randomize_platform=True, max_steps=20, auto_reset=True, dtype=torch.float32)
    big = dd.BatchedDroneEnv(n, device=DEV, **kw); big.reset()
    out = dd.policy_rollout(big, blob, T, sample=True, want="arldo")
    L = nv.lib()
    assert L.dd_policy_rollout_grid(n, 148) == (min(-(-n // 128), 148) if -(-n // 128) <= 3 * 148 else -(-n // 512))
    for base, m in ((0, 1000), (n - 4096 - (n % 128), 4096), (n - 516, 516)):
        sub = dd.BatchedDroneEnv(m, device=DEV, env_id_base=base, **kw); sub.reset()
        o = dd.policy_rollout(sub, blob, T, sample=True, want="arldo")
        for key in ("actions", "reward", "logp", "done"):
            assert torch.equal(o[key], out[key][:, base:base + m]), (key, base, m)
        assert torch.equal(o["obs"], out["obs"][:, base:base + m]), (base, m)
    s = big.stats()
    assert s["env_steps"] == n * T and s["episodes"] == s["landed"] + s["crashed"] + s["truncated"]
    assert int((out["done"] != 0).sum().item()) == s["episodes"]


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
    assert vblob.operand_dtype == "fp16"
    emu = _emulate16(sdc, x, head=1, f16=True)
    scale = ref.std().item()
    assert v.shape == ref.shape and torch.isfinite(v).all()
    assert ((v - emu).abs() / (1 + emu.abs())).max().item() < 2e-3, ((v - emu).abs() / (1 + emu.abs())).max().item()
    assert (v - ref).abs().max().item() < 0.4 and (v - ref).pow(2).mean().sqrt().item() < 0.08, \
        ((v - ref).abs().max().item(), (v - ref).pow(2).mean().sqrt().item())       # absolute, at a value scale of ~254
    vb = dd.value_forward(dd.ValueBlob(sdc, device=DEV, operands="bf16"), x.to(DEV)).cpu()
    assert ((vb - _emulate16(sdc, x, head=1, f16=False)).abs() / (1 + emu.abs())).max().item() < 2e-3
    assert (vb - ref).abs().max().item() < 0.02 * scale
    print(f"critic fixture: max / rms value error vs eager fp32: fp16 {(v - ref).abs().max().item():.3f} / "
          f"{(v - ref).pow(2).mean().sqrt().item():.4f}, bf16 {(vb - ref).abs().max().item():.3f} / {(vb - ref).pow(2).mean().sqrt().item():.4f} (value std {scale:.0f})")
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


def test_policy_and_critic_against_eager_fp32_on_the_whole_cfg4_buffer(fixture, golden_dir):
    """the probabilities that drive the actions and the values that feed dd_gae, checked against the
    notebook's eager fp32 networks (torch on the GPU, test-only) on ALL 16.4 M observation rows of a cfg-4 rollout
    (65,536 envs x 250 steps), not on a 512-row fixture.  Reported: max / rms error, the thresholded-action flip rate
    and the largest |logit| at a flip."""
    d, sd = fixture
    c = np.load(os.path.join(golden_dir, "critic_v1.npz"))
    sdc = {k_: torch.from_numpy(c[k_]) for k_ in c.files if k_.startswith("network")}
    blob, vblob = dd.PolicyBlob(sd, device=DEV), dd.ValueBlob(sdc, device=DEV)
    n, T = 65536, 250
    env = dd.BatchedDroneEnv(n, device=DEV, seed=2, randomize_drone=True, randomize_platform=True, max_steps=250,
                             auto_reset=True, dtype=torch.float32)
    env.reset()
    out = dd.policy_rollout(env, blob, T, sample=True, want="op")
    vals = dd.value_forward(vblob, out["obs"])
    pol = reference_policy(sd).to(DEV)
    crit = reference_policy(sdc, head=1).to(DEV)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    rows = out["obs"].view(-1, 15)
    probs = out["probs"].view(-1, 3)
    v = vals.view(-1)
    perr = torch.zeros((), device=DEV); verr = torch.zeros((), device=DEV)
    psq = torch.zeros((), device=DEV, dtype=torch.float64); vsq = torch.zeros((), device=DEV, dtype=torch.float64)
    flips = torch.zeros((), device=DEV, dtype=torch.int64); flip_logit = torch.zeros((), device=DEV)
    vscale_sq = torch.zeros((), device=DEV, dtype=torch.float64)
    CH = 1 << 20
    with torch.no_grad():
        for lo in range(0, rows.shape[0], CH):
            x = rows[lo:lo + CH]
            pr = pol(x)
            dp = (probs[lo:lo + CH] - pr).abs()
            perr = torch.maximum(perr, dp.max()); psq += dp.double().pow(2).sum()
            fl = (probs[lo:lo + CH] > 0.5) != (pr > 0.5)
            flips += fl.sum()
            logit = torch.log(pr.clamp_min(1e-30)) - torch.log1p(-pr.clamp_max(1 - 1e-7))
            flip_logit = torch.maximum(flip_logit, (logit.abs() * fl).max())
            vr = crit(x).squeeze(-1)
            dv = (v[lo:lo + CH] - vr).abs()
            verr = torch.maximum(verr, dv.max()); vsq += dv.double().pow(2).sum(); vscale_sq += vr.double().pow(2).sum()
    m = rows.shape[0]
    perr, verr, flips, flip_logit = perr.item(), verr.item(), int(flips.item()), flip_logit.item()
    prms, vrms, vscale = (psq.item() / (3 * m)) ** 0.5, (vsq.item() / m) ** 0.5, (vscale_sq.item() / m) ** 0.5
    print(f"cfg-4 buffer, {m} rows: probs max err {perr:.2e} rms {prms:.2e}; threshold flips {flips} ({flips / (3 * m):.2e} per action), "
          f"max |logit| at a flip {flip_logit:.3f}; critic max err {verr:.3f} rms {vrms:.4f} at value rms {vscale:.1f}")
    assert blob.operand_dtype == "fp16" and vblob.operand_dtype == "fp16"
    assert perr <= 5e-3, perr                                # (bf16 operands, round 1: 2.4e-2 on this buffer)
    assert flip_logit < 0.02 and flips / (3 * m) < 5e-4
    assert verr <= 1.0 and vrms <= 0.1, (verr, vrms)         # absolute, at a value rms of ~250
