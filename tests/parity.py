"""Shared parity checks: the new step (CUDA path, or its g++ host build) against the float64
oracle / the reference fixtures.

STATED TOLERANCES (float32 outputs of a float64-flag-chain implementation vs the float64
reference, over horizons up to 1001 steps from injected states):

  collision / goal / timeout / done flags, outcome, episode length ... bit-exact
  player x, y [px], psi [deg], intruder x, y ........................... 1e-9 abs
  obs[0:5]  (time, heading, deviation, goal distance, goal bearing) .... 5e-7 abs
  obs[5+3i] separation / d_separation_max .............................. 5e-7 abs
  obs[6+3i] d_cpa / d_cpa_max .......................................... 2e-6 abs
  obs[7+3i] v_closing / v_closing_max .................................. 2e-6 abs
  reward ................................................ 2e-5 abs + 1e-7 * |reward|
      (the +-1000 terminal bonus makes the float32 output ulp 6e-5 on terminal steps)
  episode return (float32 accumulator) .................................. 5e-3 abs

Discontinuities of the reference itself (SURVEY Q4/Q12 and the goal-bearing wrap) are
compared wrap-aware, and steps inside an epsilon-neighbourhood of a discontinuity are
excluded AND COUNTED; the caller asserts the count stays negligible.
"""
from __future__ import annotations

import numpy as np

TOL_POS = 1e-9
TOL_OBS_BASE = 5e-7
TOL_OBS_SEP = 5e-7
TOL_OBS_CPA = 2e-6
TOL_OBS_VC = 2e-6
TOL_REWARD_ABS, TOL_REWARD_REL = 2e-5, 1e-7
TOL_RETURN = 5e-3

FLAG_MASK = 15   # collision | goal | timeout | done (bit 16 = out-of-bounds is informational)


class ParityReport:
    def __init__(self):
        self.steps = 0
        self.flag_mismatch = 0
        self.excluded = 0
        self.worst = dict(obs_base=0.0, obs_sep=0.0, obs_cpa=0.0, obs_vc=0.0, reward=0.0)

    def __repr__(self):
        return f"ParityReport(steps={self.steps}, flag_mismatch={self.flag_mismatch}, excluded={self.excluded}, worst={self.worst})"


def compare_step(rep: ParityReport, obs, reward, flags, ref_obs, ref_reward, ref_flags, alive=None):
    """Compare one step of B envs.  ``alive`` masks envs whose reference episode already ended."""
    obs = np.asarray(obs, np.float64); reward = np.asarray(reward, np.float64)
    if alive is None:
        alive = np.ones(obs.shape[0], bool)
    n = int(alive.sum())
    if n == 0:
        return
    o, ro = obs[alive], ref_obs[alive]
    r, rr = reward[alive], ref_reward[alive]
    f, rf = np.asarray(flags)[alive] & FLAG_MASK, np.asarray(ref_flags)[alive] & FLAG_MASK
    rep.steps += n
    rep.flag_mismatch += int((f != rf).sum())

    # --- wrap-aware base observation: obs[1] (psi/360) and obs[4] (bearing/360) live on a circle
    d = np.abs(o[:, :5] - ro[:, :5])
    for c in (1, 4):
        d[:, c] = np.minimum(d[:, c], 1.0 - d[:, c])
    rep.worst["obs_base"] = max(rep.worst["obs_base"], float(d.max()))
    assert d.max() <= TOL_OBS_BASE, f"base obs off by {d.max():.3e}"

    # --- per-intruder triples
    sep = np.abs(o[:, 5::3] - ro[:, 5::3])
    cpa = np.abs(o[:, 6::3] - ro[:, 6::3])
    vc = np.abs(o[:, 7::3] - ro[:, 7::3])
    # Q12: d_cpa flips sign where the relative velocity's x component crosses zero -- both values are
    # "right" within rounding of v12x; accept a sign flip there, count it.
    flip = (cpa > TOL_OBS_CPA) & (np.abs(np.abs(o[:, 6::3]) - np.abs(ro[:, 6::3])) <= TOL_OBS_CPA)
    nan_both = np.isnan(o[:, 6::3]) & np.isnan(ro[:, 6::3])
    cpa = np.where(flip | nan_both, 0.0, cpa)
    rep.excluded += int(flip.sum())
    rep.worst["obs_sep"] = max(rep.worst["obs_sep"], float(sep.max()))
    rep.worst["obs_cpa"] = max(rep.worst["obs_cpa"], float(np.nanmax(cpa)))
    rep.worst["obs_vc"] = max(rep.worst["obs_vc"], float(np.nanmax(vc)))
    assert sep.max() <= TOL_OBS_SEP, f"separation obs off by {sep.max():.3e}"
    assert np.nanmax(cpa) <= TOL_OBS_CPA, f"d_cpa obs off by {np.nanmax(cpa):.3e}"
    assert np.nanmax(vc) <= TOL_OBS_VC, f"v_closing obs off by {np.nanmax(vc):.3e}"

    # --- reward: excluded where the reference's own branch variable is within rounding of its
    # discontinuity (v_closing == 0, rewards.py:54) or d_cpa flipped (Q12)
    tol = TOL_REWARD_ABS + TOL_REWARD_REL * np.abs(rr)
    dr = np.abs(r - rr)
    near_branch = (np.abs(ro[:, 7]) < 1e-6) | flip[:, 0]
    bad = (dr > tol) & ~near_branch
    rep.excluded += int(((dr > tol) & near_branch).sum())
    rep.worst["reward"] = max(rep.worst["reward"], float(np.where(near_branch, 0.0, dr).max()))
    assert not bad.any(), f"reward off by {dr[bad].max():.3e}"


def assert_flags_exact(rep: ParityReport):
    assert rep.flag_mismatch == 0, rep
    assert rep.excluded <= max(2, rep.steps // 20000), rep


def inject_from_fixture(batch, g):
    """Load the injected initial states of a ref_rollouts fixture into a batch object."""
    batch.inject_state(g["player0"][:, [0, 1, 3]], g["traffic0"], g["steps0"], g["total0"])


def fixture_rows(g, t):
    """Reference rows available at step t of a strided fixture: (obs or None, reward, flags)."""
    stride = int(g["stride"])
    obs = g["obs_strided"][t // stride] if t % stride == 0 else None
    return obs, g["reward"][t], g["flags"][t]
