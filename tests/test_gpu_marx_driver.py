"""The drop-in boundary end to end: the UNMODIFIED reference driver (marx/src/marx.c) linked against libmarxb200.so
through integration/marx_gpu_shim.c (-Wl,--wrap), run as a user runs `marx`, on the GPU box.

  * its output directory equals what the C ABI produces for the same seed / rays (batching, running arrival time,
    tags and the ExposureTime cut survive the Marx_Photon_Type bookkeeping of the shim);
  * the bulk column writer (marxb200_write_photons) is byte-identical to the reference's own marx_write_photons
    (marxio.c:403-476) fed with the same photons (MARXB200_EGRESS=stock routes them through the stock writer);
  * the detected fraction agrees with the stock CPU `marx` (own RNG) within Poisson noise.
"""
import filecmp
import glob
import os
import re
import subprocess

import numpy as np
import pytest

import marx_b200
from marx_b200 import HISTORY, read_marx_column
from tests.golden.make_golden import COMMON, CONFIGS, write_beta_image_fits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MARX_GPU = os.path.join(ROOT, "integration", "_build", "marx_gpu")
REF = os.path.join(ROOT, "oracle", "_ref")


def run_marx(exe, outdir, args, env_extra=None, check=True):
    env = dict(os.environ, MARX_DATA_DIR=os.path.join(REF, "data"))
    env.update(env_extra or {})
    cmd = [exe, "@@" + os.path.join(REF, "par", "marx.par"), "OutputDir=" + str(outdir), "OutputVectors=#ETXYZ123DxyMPOabcdSrB"] + args
    p = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    if check:
        assert p.returncode == 0, p.stdout[-2000:]
    return p


def read_dir(d):
    cols = {}
    for f in sorted(glob.glob(os.path.join(str(d), "*.dat"))):
        name, data = read_marx_column(f)
        cols[os.path.basename(f)] = data
    return cols


needs_driver = pytest.mark.skipif(not os.path.exists(MARX_GPU), reason="integration/_build/marx_gpu not built")


@needs_driver
def test_marx_gpu_fails_loudly_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = run_marx(MARX_GPU, tmp_path / "out", COMMON + CONFIGS["c1_acis_s"]["args"] + ["NumRays=1000", "dNumRays=1000"], check=False)
    assert "no CPU fallback" in p.stdout
    assert not glob.glob(str(tmp_path / "out" / "*.dat"))


@pytest.mark.gpu
@needs_driver
@pytest.mark.parametrize("config", ["c1_acis_s", "c2_hetg_acis_s", "c3_letg_hrc_s", "c4_beta_acis_i", "c4_image_acis_i", "c1_line_acis_s", "c3_hrc_i"])
def test_marx_gpu_output_equals_c_abi(config, tmp_path):
    cfg = CONFIGS[config]
    if config == "c4_image_acis_i":
        write_beta_image_fits()           # the synthetic S-ImageFile the calibration pack was made from
    # the reference's loop always collects whole batches (marx.c:545-608): NumRays=250000 -> 3 x 100000 rays
    n_par, dn, seed = 250000, 100000, 5
    n = 3 * dn
    p = run_marx(MARX_GPU, tmp_path / "out", COMMON + cfg["args"] + ["NumRays=%d" % n_par, "dNumRays=%d" % dn, "RandomSeed=%d" % seed,
                                                                      "Verbose=1"])
    assert "marxb200: ray trace on CUDA device" in p.stdout
    got = read_dir(tmp_path / "out")
    want = []
    with marx_b200.MarxB200(config, seed=seed, max_photons=dn) as m:
        first = 0
        while first < n:
            k = min(dn, n - first)
            m.trace(first, k)
            want.append(m.download().copy())
            first += k
    ph = np.concatenate(want)
    assert len(got["energy.dat"]) == len(ph) > 0
    assert (got["tag.dat"].astype(np.uint32) == ph["tag"]).all()
    assert (got["energy.dat"] == ph["energy"].astype(np.float32)).all()
    for k, f in enumerate(("xpos.dat", "ypos.dat", "zpos.dat")):
        assert (got[f] == ph["x"][:, k].astype(np.float32)).all(), f
    for k, f in enumerate(("xcos.dat", "ycos.dat", "zcos.dat")):
        assert (got[f] == ph["p"][:, k].astype(np.float32)).all(), f
    assert (got["mirror.dat"] == ph["mirror_shell"].astype(np.int16)).all()
    assert (got["detector.dat"] == ph["ccd_num"]).all()
    assert (got["xpixel.dat"] == ph["y_pixel"]).all() and (got["ypixel.dat"] == ph["z_pixel"]).all()
    assert (got["pha.dat"] == ph["pulse_height"]).all()
    if "b_energy.dat" in got:
        assert (got["b_energy.dat"] == ph["pi"]).all()
    if "order.dat" in got:
        assert (got["order.dat"] == ph["order"]).all()
    if config == "c3_letg_hrc_s":
        assert (got["hrc_u.dat"] == ph["u_pixel"]).all() and (got["hrc_v.dat"] == ph["v_pixel"]).all()
        assert (got["hrcregion.dat"] == ph["detector_region"]).all()
        for k, f in enumerate(("ofine.dat", "ocoarse1.dat", "ocoarse2.dat", "ocoarse3.dat")):
            assert (got[f] == ph["support_orders"][:, k]).all(), f
    if "DitherModel=INTERNAL" in cfg["args"]:
        assert (got["sky_ra.dat"] == ph["dither"][:, 0]).all() and (got["sky_dec.dat"] == ph["dither"][:, 1]).all()
        assert (got["det_dy.dat"] == 0).all()
    else:
        assert "sky_ra.dat" not in got
    # TIME is monotone over the batch boundaries and continues the running sum
    t = got["time.dat"].astype(np.float64)
    assert (np.diff(t) >= 0).all()
    tot, det = re.findall(r"Total photons: (\d+), Total Photons detected: (\d+)", p.stdout)[-1]
    assert int(tot) == n and int(det) == len(ph)


@pytest.mark.gpu
@needs_driver
@pytest.mark.parametrize("config", ["c1_acis_s", "c2_hetg_acis_s", "c3_letg_hrc_s"])
def test_bulk_writer_is_byte_identical_to_the_stock_writer(config, tmp_path):
    args = COMMON + CONFIGS[config]["args"] + ["NumRays=120000", "dNumRays=50000", "RandomSeed=3"]
    run_marx(MARX_GPU, tmp_path / "bulk", args)
    run_marx(MARX_GPU, tmp_path / "stock", args, env_extra={"MARXB200_EGRESS": "stock"})
    files = sorted(os.path.basename(f) for f in glob.glob(str(tmp_path / "stock" / "*.dat")))
    assert len(files) >= 15 and files == sorted(os.path.basename(f) for f in glob.glob(str(tmp_path / "bulk" / "*.dat")))
    match, mismatch, errors = filecmp.cmpfiles(str(tmp_path / "bulk"), str(tmp_path / "stock"), files, shallow=False)
    assert not mismatch and not errors, (mismatch, errors)
    assert os.path.getsize(tmp_path / "bulk" / "energy.dat") > 32 + 4 * 1000


@pytest.mark.gpu
@needs_driver
def test_write_photons_abi_matches_download(tmp_path):
    """the same writer through the C ABI alone (no driver): two batches appended, columns vs download()"""
    mask = sum(HISTORY[k] for k in ("ENERGY", "TIME", "TAG", "DET_PIXEL", "DET_NUM", "PULSEHEIGHT", "PI", "ORDER", "MIRROR_SHELL"))
    parts, total = [], 0.0
    os.makedirs(tmp_path / "o")
    with marx_b200.MarxB200("c2_hetg_acis_s", seed=2, max_photons=1 << 18) as m:
        for b in range(2):
            m.trace(b << 18, 1 << 18)
            m.write_photons(tmp_path / "o", mask, b == 0, total)
            parts.append((m.download().copy(), total))
            total = m.counts()[2]           # the driver's accumulated total_time = running end of the arrival-time sum
    got = read_dir(tmp_path / "o")
    assert sorted(got) == ["b_energy.dat", "detector.dat", "energy.dat", "mirror.dat", "order.dat", "pha.dat", "tag.dat",
                           "time.dat", "xpixel.dat", "ypixel.dat"]
    ph = np.concatenate([p for p, _ in parts])
    assert (got["tag.dat"].astype(np.uint32) == ph["tag"]).all()
    assert (got["pha.dat"] == ph["pulse_height"]).all() and (got["xpixel.dat"] == ph["y_pixel"]).all()
    t = np.concatenate([(p["arrival_time"] + t0).astype(np.float32) for p, t0 in parts])
    assert (got["time.dat"] == t).all()


@pytest.mark.gpu
@needs_driver
def test_marx_gpu_exposure_time_cut(tmp_path):
    """ExposureTime > 0 (the marx.par default): the loop ends on the first ray at or beyond the exposure (source.c:323-334)"""
    cfg = CONFIGS["c1_acis_s"]
    common = [a for a in COMMON if not a.startswith("ExposureTime")]
    expo, dn, seed = 20000.0, 40000, 4
    p = run_marx(MARX_GPU, tmp_path / "out", common + cfg["args"] + ["ExposureTime=%g" % expo, "NumRays=100000000", "dNumRays=%d" % dn,
                                                                      "RandomSeed=%d" % seed, "Verbose=1"])
    total = int(re.findall(r"Total photons: (\d+)", p.stdout)[-1])
    got = read_dir(tmp_path / "out")
    # the same through the C ABI
    collected, t_total = 0, 0.0
    with marx_b200.MarxB200("c1_acis_s", seed=seed, max_photons=dn) as m:
        while True:
            left = expo - t_total
            if left <= 0:
                break
            m.create_photons(collected, dn)
            k = m.truncate_exposure(left)
            _, _, t_end = m.counts()
            collected += k
            t_total = t_end
            if k != dn:
                break
    assert total == collected and total % dn != 0
    t = got["time.dat"]
    assert t.max() <= expo * 1.001 and t.max() > 0.9 * expo
    # statistical check against the stock CPU driver (own RNG): detected counts within 5 sigma
    q = run_marx(os.path.join(REF, "marx"), tmp_path / "cpu", common + cfg["args"] + ["ExposureTime=%g" % expo, "NumRays=100000000",
                                                                                      "dNumRays=%d" % dn, "RandomSeed=%d" % seed, "Verbose=1"])
    n_cpu = len(read_dir(tmp_path / "cpu")["energy.dat"])
    n_gpu = len(got["energy.dat"])
    assert abs(n_cpu - n_gpu) < 5.0 * np.sqrt(n_cpu + n_gpu), (n_cpu, n_gpu)


@pytest.mark.gpu
@needs_driver
def test_rayfile_dump_and_reentry(tmp_path):
    """DumpToRayFile=yes writes the generated photons from the device list through the stock rayfile writer; a second run
    with SourceType=RAYFILE lets the stock host code read them back, injects them into HBM and traces them.  The history
    word of the file decides which stages still run (hrma.c:1171-1173 etc.)."""
    cfg = CONFIGS["c2_hetg_acis_s"]
    rays = str(tmp_path / "rays.dat")
    n, dn = 300000, 100000
    base = COMMON + cfg["args"] + ["NumRays=%d" % n, "dNumRays=%d" % dn, "RandomSeed=6", "Verbose=1", "RayFile=" + rays]
    run_marx(MARX_GPU, tmp_path / "dump", base + ["DumpToRayFile=yes"])
    assert os.path.getsize(rays) == 16 + 136 * n                      # magic + history + every generated photon
    raw = np.fromfile(rays, dtype=marx_b200.PHOTON_DTYPE, offset=16)
    assert (raw["tag"] == np.arange(n, dtype=np.uint32)).all() and (np.diff(raw["arrival_time"]) >= 0).all()
    assert ((raw["energy"] >= 0.3) & (raw["energy"] <= 8.0)).all()
    # re-entry
    args = [a for a in base if not a.startswith("SourceType=")] + ["SourceType=RAYFILE"]
    p = run_marx(MARX_GPU, tmp_path / "reentry", args)
    assert "Reflecting from HRMA [B200]" in p.stdout and "Detecting [B200]" in p.stdout
    got = read_dir(tmp_path / "reentry")
    # the same rays traced directly
    q = run_marx(MARX_GPU, tmp_path / "direct", base)
    ref = read_dir(tmp_path / "direct")
    n_a, n_b = len(got["energy.dat"]), len(ref["energy.dat"])
    assert n_a > 0 and abs(n_a - n_b) < 5.0 * np.sqrt(n_a + n_b), (n_a, n_b)
    # every event of the re-entry run is one of the file's photons, with its energy
    tags = got["tag.dat"].astype(np.int64)
    assert (np.diff(tags) > 0).all() and tags.max() < n
    assert (got["energy.dat"] == raw["energy"][tags].astype(np.float32)).all()
    # a file whose history says the mirror was already applied: the mirror stage must be skipped
    hist = np.fromfile(rays, dtype=np.uint64, count=2)
    patched = str(tmp_path / "rays_mirror_done.dat")
    data = bytearray(open(rays, "rb").read())
    data[8:16] = np.uint64(int(hist[1]) | 0x800).tobytes()
    open(patched, "wb").write(bytes(data))
    r = run_marx(MARX_GPU, tmp_path / "skip", [a for a in args if not a.startswith("RayFile=")] + ["RayFile=" + patched])
    assert "Reflecting from HRMA [B200]" not in r.stdout and "Diffracting from HETG [B200]" in r.stdout


@pytest.mark.gpu
@needs_driver
def test_user_source_is_generated_by_the_stock_host_code_and_traced_on_the_gpu(tmp_path):
    """SourceType=USER (s-user.c:160-236: a dlopen'ed generator; likewise SAOSAC and SIMPUT) has no device kernel: the unmodified
    host code produces the photons -- energies, directions, arrival times, tags, dither -- and the shim injects every batch into
    HBM (marxb200_upload_from), exactly like the RAYFILE re-entry.  The generator is the reference's own example
    (marx/doc/examples/user-source/point.c, built into oracle/_ref by oracle/ref/Makefile).
      * the STOCK CPU marx, told to dump instead of trace (DumpToRayFile=yes, marx.c:583-588), writes the very photons marx_gpu
        generates (the stages draw nothing from JDMrandom on the GPU path, the dump run skips them);
      * those records pushed through the C ABI batch by batch must give marx_gpu's event files, value for value;
      * the detected fraction agrees with the stock CPU run (own RNG in the stages) within Poisson noise."""
    user_so = os.path.join(REF, "user_point.so")
    if not os.path.exists(user_so):
        pytest.skip("oracle/_ref/user_point.so not built")
    cfg = CONFIGS["c2_hetg_acis_s"]
    n, dn, seed = 200000, 100000, 7
    base = (COMMON + [a for a in cfg["args"] if not a.startswith("SourceType=")]
            + ["SourceType=USER", "UserSourceFile=" + user_so, "NumRays=%d" % n, "dNumRays=%d" % dn, "RandomSeed=%d" % seed, "Verbose=1"])
    rays = str(tmp_path / "rays.dat")
    run_marx(os.path.join(REF, "marx"), tmp_path / "dump", base + ["DumpToRayFile=yes", "RayFile=" + rays])
    raw = np.fromfile(rays, dtype=marx_b200.PHOTON_DTYPE, offset=16)
    assert len(raw) == n and (raw["tag"] == np.arange(n, dtype=np.uint32)).all()
    p = run_marx(MARX_GPU, tmp_path / "out", base)
    assert "marxb200: ray trace on CUDA device" in p.stdout and "Reflecting from HRMA [B200]" in p.stdout
    got = read_dir(tmp_path / "out")
    want = []
    with marx_b200.MarxB200("c2_hetg_acis_s", seed=seed, max_photons=dn) as m:
        for first in range(0, n, dn):
            m.upload(raw[first:first + dn])           # the dump holds absolute arrival times (s-rayfile.c:145)
            m.mirror_reflect(); m.grating_diffract(); m.detect()
            want.append(m.download().copy())
    ph = np.concatenate(want)
    assert len(got["energy.dat"]) == len(ph) > 5000
    assert (got["tag.dat"].astype(np.uint32) == ph["tag"]).all()
    assert (got["energy.dat"] == ph["energy"].astype(np.float32)).all()
    for k, f in enumerate(("ypos.dat", "zpos.dat")):
        assert (got[f] == ph["x"][:, k + 1].astype(np.float32)).all(), f
    assert (got["mirror.dat"] == ph["mirror_shell"].astype(np.int16)).all()
    assert (got["detector.dat"] == ph["ccd_num"]).all() and (got["order.dat"] == ph["order"]).all()
    assert (got["xpixel.dat"] == ph["y_pixel"]).all() and (got["ypixel.dat"] == ph["z_pixel"]).all()
    assert (got["pha.dat"] == ph["pulse_height"]).all()
    assert (got["sky_ra.dat"] == ph["dither"][:, 0]).all() and (got["sky_dec.dat"] == ph["dither"][:, 1]).all()
    assert (np.diff(got["time.dat"].astype(np.float64)) >= 0).all()
    # statistics against the stock CPU run of the same parameters
    q = run_marx(os.path.join(REF, "marx"), tmp_path / "cpu", base)
    n_cpu, n_gpu = len(read_dir(tmp_path / "cpu")["energy.dat"]), len(ph)
    assert abs(n_cpu - n_gpu) < 5.0 * np.sqrt(n_cpu + n_gpu), (n_cpu, n_gpu)


@pytest.mark.gpu
@needs_driver
def test_marx_gpu_big_batches_background_writer_equals_synchronous_writer(tmp_path):
    """The throughput set-up of the drop-in driver: batches of 2^22 rays (the dNumRays maximum of the reference is only the range
    field of marx.par:9; integration/_build/par/marx.par raises it) and the background column-file writer.  The output
    directory must be byte-identical to the one the synchronous writer produces, and to small batches of the stock size."""
    par = os.path.join(ROOT, "integration", "_build", "par", "marx.par")
    if not os.path.exists(par):
        pytest.skip("integration/_build/par/marx.par not built")
    cfg = CONFIGS["c2_hetg_acis_s"]
    n, dn = 3 << 22, 1 << 22
    env = dict(os.environ, MARX_DATA_DIR=os.path.join(REF, "data"))
    outs = {}
    for name, threads, batch in (("async", "8", dn), ("sync", "0", dn), ("small", "3", 1 << 20)):
        d = tmp_path / name
        cmd = [MARX_GPU, "@@" + par, "OutputDir=" + str(d), "OutputVectors=#ETXYZ123DxyMPOabcdSrB"] + COMMON + cfg["args"] + [
            "NumRays=%d" % n, "dNumRays=%d" % batch, "RandomSeed=9"]
        p = subprocess.run(cmd, env=dict(env, MARXB200_WRITER_THREADS=threads), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
        assert p.returncode == 0, p.stdout[-2000:]
        outs[name] = d
    files = sorted(os.path.basename(f) for f in glob.glob(str(outs["sync"] / "*.dat")))
    assert len(files) >= 21
    for f in files:
        assert filecmp.cmp(str(outs["sync"] / f), str(outs["async"] / f), shallow=False), f
    # batches of another size: same rays, same draws; the arrival-time sum is associated differently (super-tile order inside a batch),
    # so TIME may differ in the last float bit, every other column is identical
    for f in files:
        a, b = read_marx_column(str(outs["sync"] / f))[1], read_marx_column(str(outs["small"] / f))[1]
        assert len(a) == len(b), f
        if f in ("time.dat", "sky_ra.dat", "sky_dec.dat", "sky_roll.dat", "xpos.dat", "ypos.dat", "zpos.dat", "xcos.dat", "ycos.dat", "zcos.dat",
                 "xpixel.dat", "ypixel.dat"):
            assert np.allclose(a, b, rtol=1e-5, atol=1e-6), f
        else:
            assert (a == b).mean() > 0.9999, f


def write_saosac_fits(path, n, seed=3):
    """a synthetic SAOSAC ray file (s-saosac.c:95-140: BINTABLE 'RAYTRACE', double columns RT_X ... RT_KEV, RT_WGHT, RT_TIME): rays
    behind the mirror, on the four shells' radii at the file's x = 0 plane, converging on the focus with a small scatter"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle", "ref"))
    from synth_acis_caldb import bintable_hdu, primary_hdu
    r = np.random.default_rng(seed)
    cap = 10079.77                                            # mm in front of the focus (the CAP; the source projects to it)
    radius = r.choice(np.array([600.0, 483.0, 426.0, 317.0]), n) + r.normal(0.0, 2.0, n)
    phi = r.uniform(0.0, 2 * np.pi, n)
    y, z = radius * np.cos(phi), radius * np.sin(phi)
    target = r.normal(0.0, 0.02, (n, 2))                      # mm at the focal plane
    d = np.stack([-cap * np.ones(n), target[:, 0] - y, target[:, 1] - z], axis=1)
    d /= np.linalg.norm(d, axis=1)[:, None]
    cols = [("RT_X", "D", 1, np.zeros(n)), ("RT_Y", "D", 1, y), ("RT_Z", "D", 1, z), ("RT_COSX", "D", 1, d[:, 0]), ("RT_COSY", "D", 1, d[:, 1]),
            ("RT_COSZ", "D", 1, d[:, 2]), ("RT_KEV", "D", 1, r.uniform(0.5, 6.0, n)), ("RT_WGHT", "D", 1, r.uniform(0.2, 1.0, n)),
            ("RT_TIME", "D", 1, 1000.0 + np.cumsum(r.exponential(0.3, n)))]
    with open(path, "wb") as f:
        f.write(primary_hdu() + bintable_hdu("RAYTRACE", cols, n))
    return path


@pytest.mark.gpu
@needs_driver
def test_saosac_rays_are_read_by_the_stock_host_code_and_traced_on_the_gpu(tmp_path):
    """SourceType=SAOSAC (s-saosac.c): rays of an external mirror ray trace, read by the UNMODIFIED host code (jdfits), weighted
    (JDMrandom >= RT_WGHT rejects a ray at the source, :206-211), assigned a mirror shell, and handed to the device with the mirror
    stage marked done.  Same three checks as for the USER source; a synthetic ray file stands in for a real SAOSAC product."""
    cfg = CONFIGS["c2_hetg_acis_s"]
    n, dn, seed = 120000, 60000, 11
    fits = write_saosac_fits(str(tmp_path / "saosac.fits"), n)
    base = (COMMON + [a for a in cfg["args"] if not a.startswith("SourceType=")]
            + ["SourceType=SAOSAC", "SAOSACFile=" + fits, "NumRays=%d" % n, "dNumRays=%d" % dn, "RandomSeed=%d" % seed, "Verbose=1"])
    rays = str(tmp_path / "rays.dat")
    run_marx(os.path.join(REF, "marx"), tmp_path / "dump", base + ["DumpToRayFile=yes", "RayFile=" + rays])
    raw = np.fromfile(rays, dtype=marx_b200.PHOTON_DTYPE, offset=16)
    assert 0.3 * n < len(raw) <= n                              # the weight test removed ~40 % of the rays at the source
    p = run_marx(MARX_GPU, tmp_path / "out", base)
    assert "marxb200: ray trace on CUDA device" in p.stdout and "Reflecting from HRMA [B200]" not in p.stdout
    assert "Diffracting from HETG [B200]" in p.stdout
    got = read_dir(tmp_path / "out")
    live = raw[(raw["flags"] & 0xFF) == 0]
    with marx_b200.MarxB200("c2_hetg_acis_s", seed=seed, max_photons=len(live) + 16) as m:
        m.upload(live)
        m.grating_diffract(); m.detect()
        ph = m.download().copy()
    assert len(got["energy.dat"]) == len(ph) > 2000
    assert (got["tag.dat"].astype(np.uint32) == ph["tag"]).all()
    assert (got["energy.dat"] == ph["energy"].astype(np.float32)).all()
    assert (got["mirror.dat"] == ph["mirror_shell"].astype(np.int16)).all() and len(np.unique(ph["mirror_shell"])) == 4
    assert (got["detector.dat"] == ph["ccd_num"]).all() and (got["order.dat"] == ph["order"]).all()
    assert (got["xpixel.dat"] == ph["y_pixel"]).all() and (got["ypixel.dat"] == ph["z_pixel"]).all()
    assert (got["pha.dat"] == ph["pulse_height"]).all()
    q = run_marx(os.path.join(REF, "marx"), tmp_path / "cpu", base)
    n_cpu, n_gpu = len(read_dir(tmp_path / "cpu")["energy.dat"]), len(ph)
    assert abs(n_cpu - n_gpu) < 5.0 * np.sqrt(n_cpu + n_gpu), (n_cpu, n_gpu)
