"""Histogram definitions shared by the statistical reference generator and the GPU statistics test."""
import numpy as np

BINS = {
    "energy": np.linspace(0.0, 12.0, 121),              # detected-photon energies: effective area vs E
    "pha": np.linspace(0, 4096, 257),
    "pi": np.linspace(0.0, 12.0, 121),
    "chipx": np.linspace(0, 1024, 65),
    "chipy": np.linspace(0, 1024, 65),
    "ccd": np.arange(-0.5, 10.5, 1.0),
    "order": np.arange(-11.5, 12.5, 1.0),
    "shell": np.arange(-0.5, 4.5, 1.0),
    # zeroth-order / direct image radius (mm) about the image centre: encircled-energy PSF
    "psf_r": np.concatenate([[0.0], np.geomspace(1e-3, 2.0, 48)]),
}


def summarize(cols):
    """cols: dict of per-event arrays (energy, pha, ccd, chipx, chipy, ypos, zpos, shell, order, pi).  Returns histograms."""
    out = {"n_detected": np.array(len(cols["energy"]), dtype=np.int64)}
    for k in ("energy", "pha", "pi", "chipx", "chipy", "ccd", "shell"):
        if cols.get(k) is not None:
            out["h_" + k] = np.histogram(np.asarray(cols[k], dtype=np.float64), BINS[k])[0].astype(np.int64)
    order = cols.get("order")
    if order is None:
        order = np.zeros(len(cols["energy"]), dtype=np.int64)
    out["h_order"] = np.histogram(np.asarray(order, dtype=np.float64), BINS["order"])[0].astype(np.int64)
    # PSF of the undispersed image on the aim-point chip (the chip with most zeroth-order events: S3 = ccd 7 for
    # ACIS-S, I3 for ACIS-I, the middle MCP for HRC-S): radius about the image centre (median)
    ccd = np.asarray(cols["ccd"]).astype(np.int64)
    zero = np.asarray(order) == 0
    aim = np.bincount(ccd[zero & (ccd >= 0)]).argmax() if (zero & (ccd >= 0)).any() else 7
    sel = zero & (ccd == aim)
    y, z = np.asarray(cols["ypos"], dtype=np.float64)[sel], np.asarray(cols["zpos"], dtype=np.float64)[sel]
    r = np.hypot(y - np.median(y), z - np.median(z)) if sel.any() else np.zeros(0)
    out["h_psf_r"] = np.histogram(r, BINS["psf_r"])[0].astype(np.int64)
    return out
