"""Device ops that sit right after a rollout in the PPO notebook, each one C-ABI call:

  * ``gae``                  compute_gae            (Actor_Critic_PPO.ipynb c15:L49-53), batched [T,N]
  * ``advantage_moments``    n, sum, sum of squares (K4)
  * ``normalize_advantages`` (adv - mean) / (std + 1e-8) with the UNBIASED std of torch.std
                             (Actor_Critic_PPO.ipynb c21:L105), moments all-reduced over ranks

No torch fallback: the arithmetic is in ``csrc/ppo_kernels.cu``.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _native as nv
from .distributed import allreduce_moments


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _need(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda or t.dtype != dtype or not t.is_contiguous():
        raise ValueError(f"{name} must be a contiguous CUDA tensor of dtype {dtype}")


def gae(rewards: torch.Tensor, values: torch.Tensor, dones: torch.Tensor, gamma: float = 0.99,
        lambda_: float = 0.95, want_returns: bool = False, out: Optional[torch.Tensor] = None,
        out_returns: Optional[torch.Tensor] = None, moments: Optional[torch.Tensor] = None):
    """``rewards`` fp32 [T,N], ``values`` fp32 [T+1,N] (bootstrap row last), ``dones`` uint8 [T,N]
    (non-zero = episode ended on that step).  Returns advantages [T,N] (and returns = adv + V).
    ``out`` / ``out_returns``: preallocated fp32 [T,N] result buffers.  ``moments``: float64[3] on the device;
    the scan ACCUMULATES (n, sum, sum of squares) of the advantages into it in the same pass, so
    ``normalize_advantages(adv, moments=m)`` needs no separate read of the buffer (zero it per iteration)."""
    _need(rewards, torch.float32, "rewards")
    _need(values, torch.float32, "values")
    _need(dones, torch.uint8, "dones")
    T, n = rewards.shape
    if tuple(values.shape) != (T + 1, n) or tuple(dones.shape) != (T, n):
        raise ValueError("values must be [T+1,N] and dones [T,N]")
    for o_, nm in ((out, "out"), (out_returns, "out_returns")):
        if o_ is not None:
            _need(o_, torch.float32, nm)
            if tuple(o_.shape) != (T, n):
                raise ValueError(f"{nm} must be [T,N]")
    adv = torch.empty_like(rewards) if out is None else out
    want_returns = want_returns or out_returns is not None
    ret = (torch.empty_like(rewards) if out_returns is None else out_returns) if want_returns else None
    if moments is not None and (moments.dtype != torch.float64 or moments.numel() != 3 or moments.device != rewards.device):
        raise ValueError("moments must be a float64[3] tensor on the rewards' device")
    nv.check(nv.lib().dd_gae_moments(rewards.data_ptr(), values.data_ptr(), dones.data_ptr(), adv.data_ptr(),
                                     None if ret is None else ret.data_ptr(), None if moments is None else moments.data_ptr(),
                                     float(gamma), float(lambda_), T, n, _stream(rewards)), "dd_gae_moments")
    return (adv, ret) if want_returns else adv


def discounted_returns(rewards: torch.Tensor, dones: Optional[torch.Tensor] = None, gamma: float = 0.99,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``compute_returns`` of the policy-gradient notebook (G_t = r_t + gamma * G_{t+1}, float64 accumulation,
    restarting behind every step whose ``dones`` byte is non-zero) for fp32 ``rewards`` [T,N]."""
    _need(rewards, torch.float32, "rewards")
    T, n = rewards.shape
    if dones is not None:
        _need(dones, torch.uint8, "dones")
        if tuple(dones.shape) != (T, n):
            raise ValueError("dones must be [T,N]")
    if out is None:
        out = torch.empty_like(rewards)
    else:
        _need(out, torch.float32, "out")
        if tuple(out.shape) != (T, n):
            raise ValueError("out must be [T,N]")
    nv.check(nv.lib().dd_discounted_returns(rewards.data_ptr(), None if dones is None else dones.data_ptr(),
                                            out.data_ptr(), float(gamma), T, n, _stream(rewards)), "dd_discounted_returns")
    return out


def advantage_moments(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """float64[3] on device: (n, sum x, sum x^2).  ACCUMULATES into ``out`` when given."""
    _need(x, torch.float32, "x")
    if out is None:
        out = torch.zeros(3, dtype=torch.float64, device=x.device)
    nv.check(nv.lib().dd_moments(x.data_ptr(), x.numel(), out.data_ptr(), _stream(x)), "dd_moments")
    return out


def normalize_advantages(x: torch.Tensor, eps: float = 1e-8, reduce: bool = True,
                         out: Optional[torch.Tensor] = None, moments: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Globally normalised advantages; ``reduce=True`` sums the moments over all ranks first.  ``moments``: this
    rank's (n, sum, sum of squares) when ``gae(..., moments=m)`` already accumulated them (left untouched)."""
    m = advantage_moments(x) if moments is None else moments.clone()
    if reduce:
        allreduce_moments(m)
    if out is None:
        out = torch.empty_like(x)
    nv.check(nv.lib().dd_normalize(x.data_ptr(), out.data_ptr(), m.data_ptr(), float(eps), x.numel(), _stream(x)),
             "dd_normalize")
    return out
