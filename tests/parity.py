"""Parity rules shared by the emulator (CPU) and GPU tests (BASELINE.json north_star).

On identical inputs, against the float64 oracle:
  * integer / indexing outputs are bit-exact (frame slicing, fftshift order, histogram counts);
    uint8 colormap indices are bit-exact except where the oracle's pre-floor value is within 1e-3
    of an integer (threshold ties);
  * dB rows / PSD: |delta| <= 1e-3 dB on bins at or above the row's noise floor (20th percentile of
    the oracle row); on the bins below it the linear power must agree to <= 1e-4 of the noise-floor
    power (a dB difference is meaningless for a bin that is a random deep null).
"""
import os

import numpy as np

# observed margins of every check of the current pytest session (written by tests/conftest.py to
# gpurun_out/parity_margins.json at session end; the committed copy is profiles/r02_parity_margins.json)
MARGINS = []


def _record(kind, what, **stats):
    MARGINS.append({"test": os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0], "check": kind, "what": what, **stats})


DB_TOL = 1e-3
REL_TOL = 1e-4
TIE_TOL = 1e-3


def check_db_rows(db_got, power_ref, eps=1e-12, what=""):
    """db_got: [F,N] float; power_ref: [F,N] float64 |X|^2 (same order)."""
    db_got = np.asarray(db_got, dtype=np.float64)
    with np.errstate(divide="ignore"):
        db_ref = 20 * np.log10(np.sqrt(power_ref) + eps)
    floor_db = np.percentile(db_ref, 20, axis=-1, keepdims=True)
    above = db_ref >= floor_db
    err_db = np.abs(db_got - db_ref)
    worst_above = float(err_db[above].max()) if above.any() else 0.0
    assert worst_above <= DB_TOL, f"{what}: {worst_above:.3e} dB error above the noise floor"
    floor_pw = (10 ** (floor_db / 20) - eps).clip(min=0) ** 2
    pw_got = (10 ** (db_got / 20) - eps).clip(min=0) ** 2
    rel = np.abs(pw_got - power_ref) / np.maximum(floor_pw, 1e-300)
    below = ~above
    worst_below = float(rel[below].max()) if below.any() else 0.0
    # diagnostic only: error of the bins below the floor relative to the bin itself (not a pass criterion: a deep null
    # of a float32 FFT carries the rounding noise of the whole frame)
    with np.errstate(divide="ignore", invalid="ignore"):
        rel_bin = np.abs(pw_got - power_ref) / np.maximum(power_ref, 1e-300)
    _record("db_rows", what, bins=int(db_ref.size), worst_db_above_floor=worst_above, worst_rel_of_floor_below=worst_below,
            worst_rel_of_bin_below=float(rel_bin[below].max()) if below.any() else 0.0,
            p999_rel_of_bin_below=float(np.quantile(rel_bin[below], 0.999)) if below.any() else 0.0)
    assert worst_below <= REL_TOL, f"{what}: linear error {worst_below:.3e} of the floor power below the floor"
    return worst_above, worst_below


def check_power(p_got, p_ref, what="", rel_tol=REL_TOL):
    """Welch sums / max-hold / PSD (linear power, last axis = bins), north_star rule: bins at or above the
    noise floor (20th percentile of the oracle's bins) within 1e-3 dB and `rel_tol` relative; bins below it
    within `rel_tol` of the noise-floor power."""
    p_got = np.asarray(p_got, dtype=np.float64)
    p_ref = np.asarray(p_ref, dtype=np.float64)
    if p_ref.size == 0:
        return 0.0
    floor = np.percentile(p_ref, 20, axis=-1, keepdims=True)
    above = p_ref >= floor
    rel = np.abs(p_got - p_ref) / np.maximum(np.where(above, np.abs(p_ref), floor), 1e-300)
    worst = float(rel.max())
    with np.errstate(divide="ignore", invalid="ignore"):
        ddb = np.abs(10 * np.log10(p_got) - 10 * np.log10(p_ref))
        rel_bin = np.abs(p_got - p_ref) / np.maximum(np.abs(p_ref), 1e-300)
    ddb = ddb[above & np.isfinite(ddb)]
    _record("power", what, bins=int(p_ref.size), worst_rel=worst, worst_db_above_floor=float(ddb.max()) if ddb.size else 0.0,
            worst_rel_of_bin_below=float(rel_bin[~above].max()) if (~above).any() else 0.0)
    assert worst <= rel_tol, f"{what}: relative error {worst:.3e} (of the bin above the floor, of the floor power below it)"
    if ddb.size:
        assert ddb.max() <= DB_TOL, f"{what}: {ddb.max():.3e} dB above the noise floor"
    return worst


def check_u8(q_got, db_ref, vmin, vmax, what="", eps=1e-12):
    """uint8 colormap indices: bit-exact away from threshold ties.

    A tie is a bin whose oracle pre-floor value is within 1e-3 of an integer (north_star), widened,
    for consistency with the PSD tolerance, to the distance the *permitted* dB error of that bin
    (1e-3 dB at or above the row's noise floor; 1e-4 of the floor power below it) can move it.
    Mismatches outside the literal 1e-3 zone must stay below one per million bins."""
    db_ref = np.asarray(db_ref, dtype=np.float64)
    scale = 256.0 / (vmax - vmin)
    pre = (db_ref - vmin) * scale
    want = np.clip(np.nan_to_num(np.floor(pre), nan=0.0, posinf=255.0, neginf=0.0), 0, 255).astype(np.uint8)
    floor_db = np.percentile(db_ref, 20, axis=-1, keepdims=True)
    with np.errstate(over="ignore", invalid="ignore"):
        below_slack = 10 * np.log10(1.0 + REL_TOL * 10 ** ((floor_db - db_ref) / 10))
    slack_db = np.where(db_ref >= floor_db, DB_TOL, np.maximum(DB_TOL, below_slack))
    dist = np.abs(pre - np.rint(pre))
    tie_literal = dist <= TIE_TOL
    tie = dist <= np.maximum(TIE_TOL, slack_db * scale)
    diff = np.asarray(q_got) != want
    bad = diff & ~tie
    outside_literal = int((diff & ~tie_literal).sum())
    _record("u8", what, bins=int(diff.size), mismatches=int(diff.sum()), mismatches_inside_literal_tie_zone=int((diff & tie_literal).sum()),
            mismatches_outside_literal_tie_zone=outside_literal, mismatches_outside_widened_zone=int(bad.sum()),
            bins_in_literal_tie_zone=int(tie_literal.sum()), bins_in_widened_zone=int(tie.sum()),
            max_index_step=int(np.abs(np.asarray(q_got).astype(np.int16) - want.astype(np.int16)).max()) if diff.size else 0)
    assert not bad.any(), f"{what}: {int(bad.sum())} colormap indices differ away from ties"
    assert outside_literal <= max(1, diff.size // 1_000_000), f"{what}: {outside_literal} mismatches outside the 1e-3 tie zone"
    return int((diff & tie).sum()), int(tie.sum())
