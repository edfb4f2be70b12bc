#!/usr/bin/env python3
"""Synthetic stand-ins for the two ACIS calibration blobs that are absent from the
reference checkout (.MISSING_LARGE_BLOBS: caldb/acisfef.fits and
caldb/acisD1999-08-13contamN0015_marx.fits).

TEST INFRASTRUCTURE.  Both the compiled reference (oracle/_ref) and the calibration
pack consumed by the CUDA path are produced from these same files, so replay parity is
meaningful; absolute ACIS physics is not (stated next to every ACIS number).

Schemas follow the reference readers:
  FEF    : marx/libsrc/acis_fef.c:339-343 (fixed columns), :836-861 (G%d_FWHM/POS/AMPL probing),
           :194-222 (regions are multiples of 32, 1-based inclusive), :769-797 (rows grouped by REGNUM)
  CONTAM : marx/libsrc/aciscontam.c:211-233 (columns), :304-309 (EXTNAME ACIS<ccd>_CONTAM),
           :457-487 (optional fxy image + FXYBLK keyword)

Deterministic: numpy Generator seeded with 20240101.  Pure numpy FITS writer (no astropy here).
"""
import sys
import os
import numpy as np

SEED = 20240101


def _card(key, value, comment=""):
    if isinstance(value, bool):
        v = "%20s" % ("T" if value else "F")
    elif isinstance(value, (int, np.integer)):
        v = "%20d" % value
    elif isinstance(value, float):
        v = "%20s" % repr(value)
    else:
        v = "'%-8s'" % value
        v = "%-20s" % v
    s = "%-8s= %s" % (key, v)
    if comment:
        s += " / " + comment
    return ("%-80s" % s)[:80]


def _pad(b, fill=b"\0"):
    n = (-len(b)) % 2880
    return b + fill * n


def primary_hdu():
    cards = [_card("SIMPLE", True), _card("BITPIX", 8), _card("NAXIS", 0), _card("EXTEND", True), "%-80s" % "END"]
    return _pad("".join(cards).encode("ascii"), b" ")


def bintable_hdu(extname, columns, nrows, extra_keys=()):
    """columns: list of (name, tform_letter, repeat, array[nrows, repeat])"""
    widths = {"J": 4, "E": 4, "D": 8, "I": 2}
    dts = {"J": ">i4", "E": ">f4", "D": ">f8", "I": ">i2"}
    naxis1 = sum(widths[t] * r for _, t, r, _ in columns)
    cards = [
        _card("XTENSION", "BINTABLE"), _card("BITPIX", 8), _card("NAXIS", 2),
        _card("NAXIS1", naxis1), _card("NAXIS2", nrows), _card("PCOUNT", 0), _card("GCOUNT", 1),
        _card("TFIELDS", len(columns)), _card("EXTNAME", extname),
    ]
    for k, v in extra_keys:
        cards.append(_card(k, v))
    for i, (name, t, r, _) in enumerate(columns, 1):
        cards.append(_card("TTYPE%d" % i, name))
        cards.append(_card("TFORM%d" % i, "%d%s" % (r, t)))
    cards.append("%-80s" % "END")
    hdr = _pad("".join(cards).encode("ascii"), b" ")
    rec = np.dtype([(name, dts[t], (r,)) for name, t, r, _ in columns])
    data = np.zeros(nrows, dtype=rec)
    for name, t, r, arr in columns:
        data[name] = np.asarray(arr).reshape(nrows, r)
    return hdr + _pad(data.tobytes())


def make_fef(path):
    rng = np.random.default_rng(SEED)
    ngauss = 6
    energies = np.array([0.1, 0.2, 0.277, 0.4, 0.525, 0.7, 0.9, 1.1, 1.3, 1.49, 1.7, 1.85, 2.0, 2.3, 2.7,
                         3.2, 3.8, 4.5, 5.4, 6.4, 7.5, 8.6, 10.0, 12.0], dtype=np.float64)
    rows = {k: [] for k in ["CCD_ID", "CHIPX_LO", "CHIPX_HI", "CHIPY_LO", "CHIPY_HI", "REGNUM", "ENERGY", "CHANNEL"]}
    g = {("G%d_%s" % (i + 1, s)): [] for i in range(ngauss) for s in ("FWHM", "POS", "AMPL")}
    regnum = 0
    for ccd in range(10):
        bi = ccd in (5, 7)  # back-illuminated chips: broader response
        for node in range(4):
            for yb in range(8):
                regnum += 1
                gain = (4.0 + 0.15 * node + 0.02 * yb + 0.05 * ccd) * 1e-3  # keV per channel
                noise = 2.0 + 0.3 * rng.random()
                neg_region = (yb % 3 == 1)      # region with a negative-amplitude correction gaussian
                tail_region = (yb % 4 == 2)     # region with a mostly-negative-side gaussian (tail sampler)
                flip_region = (node == 3)       # amplitude changing sign between energy rows
                for ie, e in enumerate(energies):
                    chan = e / gain + 1.5
                    fano = np.sqrt(noise ** 2 + 0.115 * e * 1000 / 3.65 * (2.5 if bi else 1.0)) * 3.65e-3 / gain
                    fwhm_main = 2.3548 * fano
                    prm = [
                        (fwhm_main, chan, 1.0),
                        (fwhm_main * 2.2, chan * 0.965, 0.18 + 0.1 * yb / 8.0),
                        (fwhm_main * 1.1, max(chan - 1.739 / gain, 3.0), 0.02 if e > 1.84 else 0.0),
                        (fwhm_main * 0.6, chan * 1.01, -0.06 if neg_region else 0.0),
                        (60.0, -25.0 - ie, 0.35 if tail_region else 0.0),
                        (fwhm_main * 3.0, chan * 0.8, (0.05 if ie % 2 == 0 else -0.02) if flip_region else 0.01),
                    ]
                    rows["CCD_ID"].append(ccd)
                    rows["CHIPX_LO"].append(1 + 256 * node)
                    rows["CHIPX_HI"].append(256 * (node + 1))
                    rows["CHIPY_LO"].append(1 + 128 * yb)
                    rows["CHIPY_HI"].append(128 * (yb + 1))
                    rows["REGNUM"].append(regnum)
                    rows["ENERGY"].append(e)
                    rows["CHANNEL"].append(chan)
                    for i, (fw, pos, amp) in enumerate(prm):
                        g["G%d_FWHM" % (i + 1)].append(fw)
                        g["G%d_POS" % (i + 1)].append(pos)
                        g["G%d_AMPL" % (i + 1)].append(amp)
    n = len(rows["CCD_ID"])
    cols = [(k, "J", 1, np.array(rows[k], dtype=np.int32)) for k in
            ["CCD_ID", "CHIPX_LO", "CHIPX_HI", "CHIPY_LO", "CHIPY_HI", "REGNUM"]]
    cols += [(k, "E", 1, np.array(rows[k], dtype=np.float32)) for k in ["ENERGY", "CHANNEL"]]
    for i in range(ngauss):
        for s in ("FWHM", "POS", "AMPL"):
            k = "G%d_%s" % (i + 1, s)
            cols.append((k, "E", 1, np.array(g[k], dtype=np.float32)))
    with open(path, "wb") as f:
        f.write(primary_hdu())
        f.write(bintable_hdu("FUNCTION", cols, n))


def make_contam(path):
    rng = np.random.default_rng(SEED + 1)
    n_e, n_t, blk = 96, 8, 32
    nb = 1024 // blk
    energy = np.geomspace(0.08, 12.0, n_e)
    times = np.linspace(5.0e7, 9.5e8, n_t)  # seconds since 1998.0; TStart=2023.5 extrapolates slightly
    with open(path, "wb") as f:
        f.write(primary_hdu())
        for ccd in range(10):
            with_fxy = ccd in (0, 1, 4, 5, 6)      # the others use the analytic f(x,y) forms
            layers = 3 if ccd >= 4 else 2
            cols_data = {k: [] for k in ["component", "n_energy", "energy", "mu", "n_time", "time", "tau0", "tau1", "fxy"]}
            for layer in range(layers):
                edge = [0.284, 0.532, 0.685][layer]
                mu = 2.0e0 * (energy / 0.5) ** -2.7 * (1.0 + 3.0 * (energy >= edge)) * (0.6 + 0.2 * layer)
                tau0 = (0.02 + 0.01 * layer) * (1.0 - np.exp(-times / 3.0e8)) * (1.0 + 0.02 * ccd)
                tau1 = 0.5 * tau0 * (1.0 + 0.1 * layer)
                yy, xx = np.mgrid[0:nb, 0:nb]
                fxy = ((np.abs(yy - nb / 2 + 0.5) / (nb / 2)) ** (2.0 + layer)
                       + 0.05 * rng.random((nb, nb)) + 0.02 * xx / nb)
                cols_data["component"].append(0)
                cols_data["n_energy"].append(n_e)
                cols_data["energy"].append(energy)
                cols_data["mu"].append(mu)
                cols_data["n_time"].append(n_t)
                cols_data["time"].append(times)
                cols_data["tau0"].append(tau0)
                cols_data["tau1"].append(tau1)
                cols_data["fxy"].append(fxy.reshape(-1))
            cols = [
                ("component", "J", 1, np.array(cols_data["component"], dtype=np.int32)),
                ("n_energy", "J", 1, np.array(cols_data["n_energy"], dtype=np.int32)),
                ("energy", "E", n_e, np.array(cols_data["energy"], dtype=np.float32)),
                ("mu", "E", n_e, np.array(cols_data["mu"], dtype=np.float32)),
                ("n_time", "J", 1, np.array(cols_data["n_time"], dtype=np.int32)),
                ("time", "D", n_t, np.array(cols_data["time"], dtype=np.float64)),
                ("tau0", "E", n_t, np.array(cols_data["tau0"], dtype=np.float32)),
                ("tau1", "E", n_t, np.array(cols_data["tau1"], dtype=np.float32)),
            ]
            keys = []
            if with_fxy:
                cols.append(("fxy", "E", nb * nb, np.array(cols_data["fxy"], dtype=np.float32)))
                keys.append(("FXYBLK", blk))
            f.write(bintable_hdu("ACIS%d_CONTAM" % ccd, cols, layers, keys))


if __name__ == "__main__":
    out = sys.argv[1]
    os.makedirs(out, exist_ok=True)
    make_fef(os.path.join(out, "acisfef.fits"))
    make_contam(os.path.join(out, "acisD1999-08-13contamN0015_marx.fits"))
    print("wrote synthetic ACIS FEF + contamination files to", out)
