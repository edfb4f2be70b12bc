"""Shared comparison rules for replay parity (BASELINE.json north_star):

* FP64 geometry (x, p, time) within 1e-9 relative;
* integer assignments (flags of live rays, shell, order, ccd, pha) bit-exact;
* float32 outputs (chip pixels, PI, dither angles) within a few float ulps (they are float roundings of
  FP64 values that may differ by an FP64 ulp between glibc and CUDA libm);
* a dead ray must be dead in both, and the CUDA path reports the FIRST cause of death, which must be one
  of the bits the reference set (mx_hrma.cuh header).
"""
import numpy as np

GEOM_RTOL = 1e-9
F32_RTOL = 4e-7


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    s = np.maximum(np.abs(a), np.abs(b))
    d = np.abs(a - b)
    return np.where(s > 0, d / np.where(s > 0, s, 1), 0.0)


def vrel(a, b):
    """per-photon relative error of a 3-vector: max|a-b| / max|b| (component-wise ratios are meaningless
    for components that cancel to ~0, e.g. x on the x=0 Rowland plane or on the chip plane)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    s = np.abs(b).max(axis=-1)
    d = np.abs(a - b).max(axis=-1)
    return np.where(s > 0, d / np.where(s > 0, s, 1), d)


def compare_stage(mine, ref, stage, check_time_abs=None):
    """mine/ref: PHOTON_DTYPE arrays of the same rays after `stage` (0..3).  Returns a dict of findings."""
    out = {}
    m_alive = (mine["flags"] & 0xFF) == 0
    r_alive = (ref["flags"] & 0xFF) == 0
    out["alive_mismatch"] = int((m_alive != r_alive).sum())
    dead = ~m_alive & ~r_alive
    first_cause_ok = (ref["flags"][dead] & mine["flags"][dead] & 0xFF) == (mine["flags"][dead] & 0xFF)
    out["dead_flag_not_subset"] = int((~first_cause_ok).sum())
    both = m_alive & r_alive
    out["n_alive"] = int(both.sum())
    out["live_flags_mismatch"] = int((mine["flags"][both] != ref["flags"][both]).sum())
    a, b = mine[both], ref[both]
    out["energy_max_rel"] = float(rel(a["energy"], b["energy"]).max()) if len(a) else 0.0
    out["p_max_abs"] = float(np.abs(a["p"] - b["p"]).max()) if len(a) else 0.0
    if stage >= 1:
        out["x_max_rel"] = float(vrel(a["x"], b["x"]).max()) if len(a) else 0.0
        out["shell_mismatch"] = int((a["mirror_shell"] != b["mirror_shell"]).sum())
    if stage >= 2:
        out["order_mismatch"] = int((a["order"] != b["order"]).sum() + (a["support_orders"] != b["support_orders"]).sum())
    if stage >= 3:
        out["ccd_mismatch"] = int((a["ccd_num"] != b["ccd_num"]).sum())
        out["pha_mismatch"] = int((a["pulse_height"] != b["pulse_height"]).sum())
        out["pixel_max_rel"] = float(max(rel(a["y_pixel"], b["y_pixel"]).max(), rel(a["z_pixel"], b["z_pixel"]).max())) if len(a) else 0.0
        out["int_pixel_mismatch"] = int((np.floor(a["y_pixel"]) != np.floor(b["y_pixel"])).sum()
                                        + (np.floor(a["z_pixel"]) != np.floor(b["z_pixel"])).sum())
        out["pi_max_rel"] = float(rel(a["pi"], b["pi"]).max()) if len(a) else 0.0
        out["region_mismatch"] = int((a["detector_region"] != b["detector_region"]).sum())
        out["uv_max_rel"] = float(max(rel(a["u_pixel"], b["u_pixel"]).max(), rel(a["v_pixel"], b["v_pixel"]).max())) if len(a) else 0.0
    # dither angles are float roundings of FP64 values (dither.c:173-175): a 1e-16 difference in the arrival-time
    # sum can flip the last float bit, so they are compared to a float ulp, not exactly
    # (absolute 1e-11 rad ~ one float ulp at the 16 arcsec dither amplitude, plus a relative float ulp for the roll)
    da, db = a["dither"][:, :3].astype(np.float64), b["dither"][:, :3].astype(np.float64)
    out["dither_excess"] = float((np.abs(da - db) - F32_RTOL * np.abs(db)).max()) if len(a) else 0.0
    return out


def assert_stage_ok(f, stage):
    assert f["alive_mismatch"] == 0, f
    assert f["dead_flag_not_subset"] == 0, f
    assert f["live_flags_mismatch"] == 0, f
    assert f["energy_max_rel"] == 0.0, f
    assert f["p_max_abs"] <= GEOM_RTOL, f
    if stage >= 1:
        assert f["x_max_rel"] <= GEOM_RTOL, f
        assert f["shell_mismatch"] == 0, f
    if stage >= 2:
        assert f["order_mismatch"] == 0, f
    if stage >= 3:
        assert f["ccd_mismatch"] == 0 and f["pha_mismatch"] == 0 and f["int_pixel_mismatch"] == 0, f
        assert f["pixel_max_rel"] <= F32_RTOL and f["pi_max_rel"] <= F32_RTOL and f["uv_max_rel"] <= F32_RTOL, f
        assert f["region_mismatch"] == 0, f
    assert f["dither_excess"] <= 1e-11, f
