"""Deterministic parity corpus (SURVEY.md 8c "Randomised parity corpus").

Inputs are derived from a splitmix64-style integer hash written in plain numpy
integer ops, so they are identical on every numpy version and need not be
stored: only the reference's OUTPUTS are committed as golden files.
"""
import numpy as np

N_CORPUS = 4096
T_CORPUS = 250
N_TRAJ = 32          # envs whose full per-step trajectory is stored


def _mix(z):
    z = z.astype(np.uint64)
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def _u01(stream, a, b):
    """uniform [0,1) doubles from hash(stream, a, b)."""
    a = np.asarray(a, np.uint64)
    b = np.asarray(b, np.uint64)
    with np.errstate(over="ignore"):
        h = _mix(_mix(a * np.uint64(0x100000001B3) + np.uint64(stream)) ^ (b + np.uint64(0x632BE59BD9B4E019)))
    return (h >> np.uint64(11)).astype(np.float64) / float(1 << 53)


def corpus_spawns(n=N_CORPUS):
    """integer spawns from the reference's four ranges (game_engine.py:66-83)."""
    i = np.arange(n)
    x = 100 + np.floor(_u01(1, i, 0) * 601).astype(np.int64)     # [100, 700]
    y = 50 + np.floor(_u01(2, i, 0) * 201).astype(np.int64)      # [50, 250]
    px = 100 + np.floor(_u01(3, i, 0) * 600).astype(np.int64)    # [100, 699]
    py = 100 + np.floor(_u01(4, i, 0) * 450).astype(np.int64)    # [100, 549]
    return x, y, px, py


def corpus_actions(T=T_CORPUS, n=N_CORPUS):
    """[T, n] uint8, bit0 main / bit1 left / bit2 right.
    even env ids: Bernoulli(0.5) x3; odd ids: Bernoulli(0.5, 0.1, 0.1)."""
    t = np.arange(T)[:, None]
    i = np.arange(n)[None, :]
    odd = (i & 1).astype(bool)
    p_side = np.where(odd, 0.1, 0.5)
    main = _u01(11, i, t) < 0.5
    left = _u01(12, i, t) < p_side
    right = _u01(13, i, t) < p_side
    return (main.astype(np.uint8) | (left.astype(np.uint8) << 1) | (right.astype(np.uint8) << 2))
