"""Client-side reward shaping of the PPO notebook, restated (ORACLE, test-only).

Follows ``calc_reward(state, prev_state)`` of /root/reference/Actor_Critic_PPO.ipynb (code cell 7,
lines 2-101; the unused ``time_penalty`` / ``calc_velocity_alignment`` terms of cells 6-7 do not
enter the total) and the time-out rule of ``collect_episodes_ppo`` (cell 16, lines 89-93: at
``step_count >= max_steps`` the episode is forced done and ``reward -= 500`` unless landed).  Note the
notebook's bookkeeping (cell 16, lines 71-72, 101-102): the ``prev_state`` handed to ``calc_reward`` at
step k is the state observed BEFORE step k-1, i.e. two states behind ``next_state``; it is ``None`` on
the first step of an episode.  All quantities are the normalised observation fields.

Parity status: pinned -- ``tests/golden/make_shaping_golden.py`` executes the notebook's own cell
source and records its outputs; ``tests/test_oracle_golden.py`` checks this port against them bit
for bit.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

# observation column indices (game_engine.py:160-177 / Actor_Critic_PPO.ipynb c10:L3-19)
VX, VY, ANGLE, FUEL, DIST, DX, DY, SPEED, LANDED, CRASHED = 2, 3, 4, 6, 9, 10, 11, 12, 13, 14


def shaped_reward(obs: Sequence[float], prev_dist: Optional[float]) -> float:
    """``calc_reward(state, prev_state)['total']`` with ``obs`` the 15 normalised fields of ``state``
    and ``prev_dist = prev_state.distance_to_platform`` (``None`` when there is no prev_state)."""
    total = 0
    total += -0.5
    dist = obs[DIST]
    speed = obs[SPEED]
    r_distance = 0
    r_hover = 0
    if prev_dist is not None:
        delta = prev_dist - dist
        if dist > 1e-6:
            toward = (obs[VX] * obs[DX] + obs[VY] * obs[DY]) / dist
        else:
            toward = 0.0
        if speed >= 0.15 and toward > 0.1 and dist > 0.065:
            r_distance = float(min(max(delta * 1000 * (1.0 + speed * 2.0), -2), 5))
        elif delta < -0.001:
            r_distance = -2.0 * abs(delta) * 1000
        elif speed < 0.05:
            r_hover = -1.0
        elif speed < 0.15:
            r_hover = -0.3
    total += r_distance
    total += r_hover
    excess = abs(obs[ANGLE]) - (((0.20 - 0.111) * dist) + 0.111)
    total += -max(excess, 0)
    if dist < 1:
        total += -2 * max(speed - 0.1, 0)
    else:
        total += -1 * max(speed - 0.6, 0)
    if obs[DY] > 0:
        total += 0
    else:
        total += obs[DY] * 4.0
    terminal = 0
    if obs[LANDED]:
        terminal = 800.0 + obs[FUEL] * 100.0
    elif obs[CRASHED]:
        terminal = -200.0
        if dist > 0.3:
            terminal -= 100.0
    total += terminal
    return total


def shaped_reward_pg(obs: Sequence[float]) -> float:
    """``calc_reward(state)['total']`` of /root/reference/Policy_Gradients.ipynb (code cell 6; helpers
    ``calc_velocity_alignment`` of code cell 5 and ``inverse_quadratic`` / ``scaled_shifted_negative_sigmoid`` of
    rl_helpers/scalers.py:14-15,20-21), in the notebook's statement order.  Stateless: no prev_state.
    Note the notebook's sign convention in calc_velocity_alignment: ``optimal_dx = -state.dx_to_platform``."""
    total = 0
    dist = obs[DIST]
    speed = obs[SPEED]
    # time penalty: -inverse_quadratic(dist, decay=50, scaler=1-0.3) - 0.3
    total += -((1 - 0.3) * (1 / (1 + (50 * (dist ** 2))))) - 0.3
    # velocity alignment (cosine of velocity vs "optimal" direction)
    odx, ody = -obs[DX], -obs[DY]
    onorm = math.sqrt(odx ** 2 + ody ** 2)
    if onorm < 1e-6:
        va = 1.0
    else:
        odx /= onorm
        ody /= onorm
        if speed < 1e-6:
            va = 0.0
        else:
            va = (obs[VX] / speed) * odx + (obs[VY] / speed) * ody
    r_distance = 0
    r_align = 0
    if dist > 0.065 and obs[DY] > 0:
        r_distance = int(va > 0) * speed * (4.5 * (1 / (1 + math.exp(10 * (dist - 0.5)))))
        if va > 0:
            r_align = 0.5
    total += r_distance
    total += r_align
    excess = abs(obs[ANGLE]) - (((0.20 - 0.111) * dist) + 0.111)
    total += -max(excess, 0)
    if dist < 1:
        total += -2 * max(speed - 0.1, 0)
    else:
        total += -1 * max(speed - 0.4, 0)
    if obs[DY] > 0:
        total += 0
    else:
        total += obs[DY] * 4.0
    terminal = 0
    if obs[LANDED]:
        terminal = 500.0 + obs[FUEL] * 100.0
    elif obs[CRASHED]:
        terminal = -200.0
        if dist > 0.3:
            terminal -= 100.0
    total += terminal
    return total


class EpisodeShaper:
    """The notebook's per-game bookkeeping: feed the state the policy saw and the state after the
    step; returns the training reward of that step (with the time-out penalty).  ``variant``: 'ppo'
    (Actor_Critic_PPO / Actor_Critic_Basic: calc_reward(state, prev_state)) or 'pg' (Policy_Gradients:
    calc_reward(state); the same -500 time-out rule, Policy_Gradients.ipynb collect_episodes)."""

    def __init__(self, max_steps: int = 0, variant: str = "ppo"):
        self.max_steps = int(max_steps)
        self.variant = variant
        self.reset()

    def reset(self) -> None:
        self.prev_dist = None        # prev_game_states[g] = None
        self.count = 0

    def step(self, current_obs: Sequence[float], next_obs: Sequence[float]):
        r = shaped_reward(next_obs, self.prev_dist) if self.variant == "ppo" else shaped_reward_pg(next_obs)
        self.count += 1
        timed_out = self.max_steps > 0 and self.count >= self.max_steps
        if timed_out and not next_obs[LANDED]:
            r -= 500
        self.prev_dist = current_obs[DIST]     # prev_game_states[g] = current_state
        return r, timed_out
