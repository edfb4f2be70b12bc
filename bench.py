#!/usr/bin/env python3
"""bench.py -- rays/sec of the MARX ray-trace hot path (HRMA -> HETG -> ACIS-S with dither).

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (one process per GPU)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU implementation, all host cores

A "step" is one batch of 2**24 generated rays per GPU of BASELINE.json configs[1]
(point source, flat 0.3-8 keV, HETG + ACIS-S, INTERNAL dither): marx_create_photons ->
marx_mirror_reflect -> marx_grating_diffract -> marx_detect, i.e. the body of marx.c:545-608 without the
file output.  The photon SoA of one batch is ~1.7 GB, far larger than the 126 MB L2, so no L2 flush is
needed between steps (stated in config.l2).  ACIS calibration is synthetic (the FEF/contamination blobs
are absent from the reference checkout); both arms load the same files.

Prints ONE JSON line (rank 0).  `value` = generated rays/s with everything resident in HBM;
`e2e` = the same through the C ABI with the event list copied to pinned host memory every step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rays/sec HRMA+HETG+ACIS-S"
UNIT = "rays/s"
WORKLOAD = ("BASELINE.json configs[1]: point source, flat 0.3-8 keV, HETG+ACIS-S, INTERNAL dither; "
            "one step = one batch of 2^24 generated rays per GPU (synthetic ACIS FEF/contamination)")
REF_ARGS = ["ExposureTime=0", "Verbose=0", "SourceFlux=0.003", "TStart=2023.5", "SourceType=POINT",
            "SpectrumType=FLAT", "MinEnergy=0.3", "MaxEnergy=8.0", "GratingType=HETG", "DetectorType=ACIS-S",
            "DitherModel=INTERNAL", "dNumRays=1000000"]

# algorithmic bytes per input ray of each staged kernel (SURVEY.md 8d / DESIGN.md "roofline")
BYTES_PER_RAY = {"K0": 56.0, "K1": 89.0, "K2": 114.0, "K3": 147.0}
# FP64 flop-equivalents per input ray (SURVEY.md 8d contract figures)
FLOPEQ_PER_RAY = {"K0": 530.0, "K1": 860.0, "K2": 700.0, "K3": 850.0}


# ------------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def ref_paths():
    ref = os.path.join(ROOT, "oracle", "_ref")
    return ref, os.path.join(ref, "marx_trace_bench"), os.path.join(ref, "par", "marx.par"), os.path.join(ref, "data")


def run_reference_sample(nrays_per_proc, nprocs, seed0=1, exe_name="marx_trace_bench"):
    """nprocs concurrent single-threaded reference processes (the reference's own multi-core mode:
    N independent `marx` runs); returns (total rays, wall seconds of the trace loops)."""
    ref, exe, par, data = ref_paths()
    exe = os.path.join(ref, exe_name)
    if not os.path.exists(exe):
        raise RuntimeError("oracle/_ref/%s not built (run __graft_entry__.build() where /root/reference exists)" % exe_name)
    env = dict(os.environ, MARX_DATA_DIR=data)
    t0 = time.time()
    procs = [subprocess.Popen([exe, str(nrays_per_proc), "@@" + par, "RandomSeed=%d" % (seed0 + k)] + REF_ARGS,
                              stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, env=env) for k in range(nprocs)]
    outs = [p.communicate()[0] for p in procs]
    wall = time.time() - t0
    rays, secs = 0, 0.0
    for p, o in zip(procs, outs):
        if p.returncode != 0:
            raise RuntimeError("reference process failed")
        j = json.loads(o.decode().strip().splitlines()[-1])
        rays += j["rays"]
        secs = max(secs, j["seconds"])
    return rays, secs, wall


def write_marx_column(path, name, kind, data):
    """one column file of a MARX output directory (marxio.c:151-205: 32-byte header + big-endian data); bench harness only, for the
    directory the STOCK marxpileup is timed on"""
    import numpy as np
    dt = {"E": ">f4", "A": "i1", "I": ">i2", "J": ">i4"}[kind]
    hdr = bytearray(32)
    hdr[0:4] = bytes([0x83, 0x13, 0x89, 0x8D])
    hdr[4] = ord(kind)
    nm = name.encode()[:15]
    hdr[5:5 + len(nm)] = nm
    hdr[20:24] = int(len(data)).to_bytes(4, "big")
    hdr[24:28] = (1).to_bytes(4, "big")
    with open(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(np.ascontiguousarray(data).astype(dt).tobytes())


def scratch_dir(name):
    import shutil
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else "/tmp"
    d = os.path.join(base, "marxb200_bench_%d_%s" % (os.getpid(), name))
    shutil.rmtree(d, ignore_errors=True)
    return d


def level1_reference_baseline(args, n_rays=1 << 24):
    """the STOCK marx2fits (oracle/_ref/marx2fits, gcc -O2, one core) on the event files of a C2 simulation of the bench's batch size,
    written by the drop-in driver (integration/_build/marx_gpu = the unmodified marx.c on the CUDA path) to tmpfs"""
    import shutil
    ref, _, par, data = ref_paths()
    marx_gpu = os.path.join(ROOT, "integration", "_build", "marx_gpu")
    m2f = os.path.join(ref, "marx2fits")
    if not (os.path.exists(marx_gpu) and os.path.exists(m2f)):
        return {"value": None, "kind": "reference", "sample": "unavailable: integration/_build/marx_gpu or oracle/_ref/marx2fits not built"}
    d = scratch_dir("l1")
    env = dict(os.environ, MARX_DATA_DIR=data, USER=os.environ.get("USER", "marx"))
    try:
        t0 = time.time()
        p = subprocess.run([marx_gpu, "@@" + par, "OutputDir=" + d, "NumRays=%d" % n_rays, "RandomSeed=%d" % args.seed] + REF_ARGS,
                           env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        t_marx = time.time() - t0
        if p.returncode != 0:
            return {"value": None, "kind": "reference", "sample": "unavailable: marx_gpu failed: %s" % p.stdout[-200:]}
        rows = (os.path.getsize(os.path.join(d, "pha.dat")) - 32) // 2
        fits = os.path.join(d, "evt.fits")
        t0 = time.time()
        p = subprocess.run([m2f, "--pixadj=edser", d, fits], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        dt = time.time() - t0
        if p.returncode != 0:
            return {"value": None, "kind": "reference", "sample": "unavailable: marx2fits failed: %s" % p.stdout[-200:]}
        return {"value": rows / dt, "unit": "events/s", "cores": 1, "kind": "reference", "cpu": cpu_model(),
                "sample": "%d events of a %d-ray C2 simulation: stock marx2fits --pixadj=edser (reads 20 column files from tmpfs, writes "
                          "the Level-1 FITS file to tmpfs), %.2f s; the directory came from marx_gpu in %.2f s" % (rows, n_rays, dt, t_marx)}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def pileup_reference_baseline(cols, alpha, ft, n_ref):
    """the STOCK marxpileup (oracle/_ref/marxpileup, gcc -O2, one core) on the first n_ref events of the list, as a MARX output
    directory on tmpfs; -> (cpu_baseline object, its output rows)"""
    import shutil
    ref, _, par, data = ref_paths()
    exe = os.path.join(ref, "marxpileup")
    if not os.path.exists(exe):
        return {"value": None, "kind": "reference", "sample": "unavailable: oracle/_ref/marxpileup not built"}
    d = scratch_dir("pu")
    os.makedirs(d)
    files = {"ccd": ("detector.dat", "CCDID", "A"), "x": ("xpixel.dat", "CHIPX", "E"), "y": ("ypixel.dat", "CHIPY", "E"),
             "t": ("time.dat", "TIME", "E"), "benergy": ("b_energy.dat", "B_ENERGY", "E"), "sky_ra": ("sky_ra.dat", "RA", "E"),
             "sky_dec": ("sky_dec.dat", "DEC", "E"), "sky_roll": ("sky_roll.dat", "ROLL", "E"), "det_dy": ("det_dy.dat", "DET_DY", "E"),
             "det_dz": ("det_dz.dat", "DET_DZ", "E"), "det_theta": ("det_theta.dat", "DET_THETA", "E")}
    try:
        for k, (f, nm, kind) in files.items():
            write_marx_column(os.path.join(d, f), nm, kind, cols[k][:n_ref])
        write_marx_column(os.path.join(d, "energy.dat"), "ENERGY", "E", cols["benergy"][:n_ref])
        shutil.copy(par, os.path.join(d, "marx.par"))            # DetectorType=ACIS-S is the file's default
        open(os.path.join(d, "obs.par"), "w").write("# synthetic list\n")
        env = dict(os.environ, MARX_DATA_DIR=data, USER=os.environ.get("USER", "marx"))
        t0 = time.time()
        p = subprocess.run([exe, "@@" + os.path.join(ref, "par", "marxpileup.par"), "MarxOutputDir=" + d, "Verbose=0", "Alpha=%r" % alpha,
                            "FrameTime=%r" % ft, "FrameTransferTime=0.0"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           text=True, timeout=600)
        dt = time.time() - t0
        if p.returncode != 0:
            return {"value": None, "kind": "reference", "sample": "unavailable: marxpileup failed: %s" % p.stdout[-200:]}
        rows = (os.path.getsize(os.path.join(d, "pileup", "pha.dat")) - 32) // 2
        return {"value": n_ref / dt, "unit": "events/s", "cores": 1, "kind": "reference", "cpu": cpu_model(), "rows": rows,
                "sample": "%d events (11 column files on tmpfs in, 14 out), stock marxpileup incl. its FEF read, %.2f s" % (n_ref, dt)}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def driver_leg(args, n_rays=127 << 24, dn=1 << 24, ref_rays=2000000):
    """What a MARX user gets: the UNMODIFIED reference driver on the CUDA path (integration/_build/marx_gpu = marx/src/marx.c linked
    against libmarxb200.so with -Wl,--wrap) run as `marx`, 127 x 2^24 rays of C2 in batches of 2^24 (dNumRays: the reference's maximum of
    10^6 is only the range field of marx/par/marx.par:9, raised in integration/_build/par/marx.par), event files written to tmpfs by
    the background writer; wall clock of the whole process, CUDA start-up and calibration-file reading included.  Beside it the stock
    CPU `marx` (oracle/_ref/marx) on a bounded sample with the same arguments."""
    import shutil
    ref, _, par_stock, data = ref_paths()
    marx_gpu = os.path.join(ROOT, "integration", "_build", "marx_gpu")
    par = os.path.join(ROOT, "integration", "_build", "par", "marx.par")
    if not (os.path.exists(marx_gpu) and os.path.exists(par)):
        return {"unavailable": "integration/_build/marx_gpu not built"}
    env = dict(os.environ, MARX_DATA_DIR=data, USER=os.environ.get("USER", "marx"), MARXB200_TIMING="1")
    # 127 batches (just under the int NumRays of marx.c:86-87) write 11 GB of event files: fall back to 2^30 / 2^28 rays on a small tmpfs
    free = shutil.disk_usage(os.path.dirname(scratch_dir("probe"))).free
    if free < (24 << 30):
        n_rays = (1 << 30) if free >= (12 << 30) else (1 << 28)
    common = [a for a in REF_ARGS if not a.startswith("dNumRays")]
    d = scratch_dir("drv")
    out = {}
    try:
        t0 = time.time()
        p = subprocess.run([marx_gpu, "@@" + par, "OutputDir=" + d, "NumRays=%d" % n_rays, "dNumRays=%d" % dn, "RandomSeed=%d" % args.seed] + common,
                           env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
        wall = time.time() - t0
        if p.returncode != 0:
            return {"unavailable": "marx_gpu failed: %s" % p.stdout[-300:]}
        rows = (os.path.getsize(os.path.join(d, "pha.dat")) - 32) // 2
        nbytes = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d) if f.endswith(".dat"))
        timing = [ln for ln in p.stdout.splitlines() if "host seconds" in ln]
        out = {"rays": n_rays, "rays_per_batch": dn, "wall_s": wall, "rays_per_s": n_rays / wall, "events": rows, "event_file_bytes": nbytes,
               "writer_threads": int(os.environ.get("MARXB200_WRITER_THREADS", "8")), "breakdown": timing[-1].split("marxb200: ")[-1] if timing else None}
        if timing:
            import re
            f = {k: float(v) for k, v in re.findall(r"(init\+upload|CUDA context|create_photons|stages|write_photons) ([0-9.]+)", timing[-1])}
            if {"create_photons", "stages", "write_photons"} <= set(f):
                loop = f["create_photons"] + f["stages"] + f["write_photons"]
                out["loop_s"] = loop
                out["loop_rays_per_s"] = n_rays / loop
                out["cuda_context_s"] = f.get("CUDA context")
                out["note"] = ("wall_s is the whole process: CUDA context creation on this box (cuda_context_s) dominates a run the int NumRays of "
                               "marx.c:86-87 caps at 2^31 rays; loop_s = host seconds inside the wrapped per-batch calls (create, stages, write) = "
                               "the rate of the collection loop itself, files written to tmpfs by the background writer")
    finally:
        shutil.rmtree(d, ignore_errors=True)
    if not args.no_cpu_baseline and os.path.exists(os.path.join(ref, "marx")):
        d = scratch_dir("drv_ref")
        try:
            t0 = time.time()
            p = subprocess.run([os.path.join(ref, "marx"), "@@" + par_stock, "OutputDir=" + d, "NumRays=%d" % ref_rays, "RandomSeed=%d" % args.seed] + REF_ARGS,
                               env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
            dt = time.time() - t0
            if p.returncode == 0:
                out["cpu_baseline"] = {"value": ref_rays / dt, "unit": UNIT, "cores": 1, "kind": "reference", "cpu": cpu_model(),
                                       "sample": "stock marx, %d rays, same arguments, files to tmpfs, %.1f s wall" % (ref_rays, dt)}
        finally:
            shutil.rmtree(d, ignore_errors=True)
    return out


def level1_leg(m, stream, args, reps=20):
    """marxb200_level1_transform on the event list the last traced batch left on the device (EDSER sub-pixel mode, the
    marx2fits default), timed with the library's CUDA-event marks; beside it the plain-C restatement of marx2fits' loop
    (oracle/level1_oracle.c, one core) on the same events -- the `cpu_baseline` of this leg."""
    import numpy as np
    import torch
    from marx_b200.level1 import Level1Desc
    z = np.load(os.path.join(ROOT, "tests", "golden", "level1_acis_s_hetg_edser.npz"))
    desc = {k[5:]: z[k] for k in z.files if k.startswith("desc.")}
    with torch.cuda.stream(stream):
        m.set_level1(Level1Desc.from_dict(desc))
        events = int(m.counts()[1])
        for _ in range(3):
            m.level1_transform(0.0)
        m.set_profiling(True)
        m.kernel_ms()
        for _ in range(reps):
            m.level1_transform(0.0)
        k = m.kernel_ms()["level1"]
        m.set_profiling(False)
    ms = k[0] / max(k[1], 1)
    # algorithmic bytes per event: R time 8 + chip pixels 8 + PI energy 4 + pha 2 + ccd 1 + aspect 12 = 35, W 5 f64 + 6 i32 + f32 +
    # 8 i16 + keep = 85
    out = {"events": events, "ms": ms, "events_per_s": events / (ms * 1e-3), "launches_per_transform": 4,
           "hbm_gbs_at_120B_per_event": 120.0 * events / (ms * 1e-3) / 1e9,
           "workload": "events of one 2^24-ray C2 batch, ACIS-S, EDSER sub-pixel mode; 3 kernels + state hand-over"}
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = level1_reference_baseline(args)
    return out


def pileup_leg(m, stream, args, n=1 << 22, reps=5):
    """marxb200_pileup_run (marxpileup's frame loop) on a synthetic bright-source event list: ~32 events per
    3.241 s exposure frame on a spot of sigma 3 pixels.  `ms` = the device kernel(s) (CUDA events inside the library), `e2e_ms` = the
    whole call with PINNED host columns in and out.  Beside it the STOCK marxpileup on the same list (one core, files on tmpfs) and
    the pinned plain-C oracle as the bit-for-bit check."""
    import numpy as np
    import torch
    r = np.random.default_rng(args.seed)
    alpha, ft, rate = 0.5, 3.241, 10.0
    src = {"ccd": np.full(n, 7, np.int8), "t": np.cumsum(r.exponential(1.0 / rate, n)).astype(np.float32),
           "x": (512.0 + r.normal(0.0, 3.0, n)).astype(np.float32), "y": (300.0 + r.normal(0.0, 3.0, n)).astype(np.float32),
           "benergy": r.uniform(0.4, 7.0, n).astype(np.float32)}
    for k in ("sky_ra", "sky_dec", "sky_roll", "det_dy", "det_dz", "det_theta"):
        src[k] = r.normal(0.0, 1e-3, n).astype(np.float32)

    def pinned_like(a, rows=None):
        t = torch.empty(len(a) if rows is None else rows, dtype=torch.from_numpy(a[:1]).dtype, pin_memory=True)
        return t.numpy()
    cols = {}
    for k, v in src.items():
        cols[k] = pinned_like(v)
        cols[k][:] = v
    out = {"ccd": pinned_like(src["ccd"]), "frame": pinned_like(np.zeros(1, np.int32), n), "nphotons": pinned_like(np.zeros(1, np.int16), n),
           "pha": pinned_like(np.zeros(1, np.int16), n)}
    for k in ("x", "y", "t", "benergy", "sky_ra", "sky_dec", "sky_roll", "det_dy", "det_dz", "det_theta"):
        out[k] = pinned_like(src["x"])
    with torch.cuda.stream(stream):
        before = m.launch_count()
        got, _ = m.pileup(cols, alpha, ft, args.seed, out=out)
        launches = m.launch_count() - before
        ms, wall = [], []
        for _ in range(reps):
            t0 = time.time()
            got, k = m.pileup(cols, alpha, ft, args.seed, out=out)
            wall.append((time.time() - t0) * 1e3)
            ms.append(k)
    ms, wall, rows = float(np.median(ms)), float(np.median(wall)), len(got["t"])
    got = {k: v.copy() for k, v in got.items()}
    # algorithmic bytes: R 41 per event (ccd 1, pixel/time/energy 16, aspect 24), W 49 per output row
    alg = 41.0 * n + 49.0 * rows
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        pass
    res = {"events": n, "rows": rows, "ms": ms, "events_per_s": n / (ms * 1e-3), "e2e_ms": wall, "e2e_events_per_s": n / (wall * 1e-3),
           "launches_per_call": launches, "hbm_gbs_algorithmic": alg / (ms * 1e-3) / 1e9, "hbm_frac": alg / (ms * 1e-3) / 1e9 / peak,
           "h2d_bytes": 41 * n, "d2h_bytes": 49 * rows,
           "workload": "synthetic list, %d events, %.0f events/s, frame %.3f s, alpha %.1f, one chip; host columns in pinned memory" % (n, rate, ft, alpha)}
    if not args.no_cpu_baseline:
        try:
            res["cpu_baseline"] = pileup_reference_baseline(src, alpha, ft, n)
            if res["cpu_baseline"].get("rows") is not None:
                res["rows_within_1pct_of_stock"] = bool(abs(res["cpu_baseline"]["rows"] - rows) < 0.01 * rows)      # the stock program draws from its own generator
        except Exception as e:  # noqa: BLE001
            res["cpu_baseline"] = {"value": None, "kind": "reference", "sample": "unavailable: %s" % str(e)[:120]}
        try:
            from tests import pileup_lib
            t0 = time.time()
            ref = pileup_lib.oracle_pileup(src, ["Alpha=%r" % alpha, "FrameTime=%r" % ft, "FrameTransferTime=0.0"], args.calpack, args.seed)
            dt = time.time() - t0
            res["parity"] = bool(all(got[k].tobytes() == ref[k].tobytes() for k in ref))
            res["oracle_port"] = {"value": n / dt, "unit": "events/s", "cores": 1, "kind": "port",
                                  "sample": "%d events, oracle/pileup_oracle.c (gcc -O2, arrays in memory), %.2f s" % (n, dt)}
        except Exception as e:  # noqa: BLE001
            res["parity"] = None
            res["oracle_port"] = {"value": None, "kind": "port", "sample": "unavailable: %s" % str(e)[:120]}
    return res


def bind_to_gpu_numa_node(index):
    """Pin this rank's host threads to the CPUs local to its GPU (sysfs local_cpulist of the GPU's PCI device), so that
    the pinned egress buffers are first-touched on that NUMA node and the D2H copies of 8 ranks do not all cross the
    socket interconnect.  Returns a short description for the JSON line (None if the topology is not exposed)."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]                      # sysfs uses a 4-digit PCI domain
        base = "/sys/bus/pci/devices/" + bus
        cpus = open(base + "/local_cpulist").read().strip()
        node = open(base + "/numa_node").read().strip()
        ids = set()
        for part in cpus.split(","):
            if "-" in part:
                a, b = part.split("-")
                ids.update(range(int(a), int(b) + 1))
            elif part:
                ids.add(int(part))
        ids &= set(os.sched_getaffinity(0))
        if not ids:
            return None
        os.sched_setaffinity(0, ids)
        return "gpu %d: numa node %s, %d cpus" % (index, node, len(ids))
    except Exception:  # noqa: BLE001
        return None


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def d2h_probe(m, rank, world, dev, barrier, nbytes=128 << 20, reps=6):
    """marxb200_probe_d2h: pinned device-to-host copies of 128 MB; every rank alone in turn, then all ranks together."""
    import torch
    import torch.distributed as dist
    alone = 0.0
    for r in range(world):
        barrier()
        if r == rank:
            alone = m.probe_d2h(nbytes, reps)
    barrier()
    together = m.probe_d2h(nbytes, reps, together=(world > 1))
    barrier()
    if world > 1:
        t = torch.tensor([alone, together], device=dev, dtype=torch.float64)
        parts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        alone_all = [float(p[0]) for p in parts]
        tog_all = [float(p[1]) for p in parts]
    else:
        alone_all, tog_all = [alone], [together]
    return {"bytes_per_copy": nbytes, "copies": reps, "alone_gbs_per_rank": alone_all, "concurrent_gbs_per_rank": tog_all,
            "concurrent_gbs_total": sum(tog_all),
            "note": "cudaHostAlloc'ed buffer, one cudaMemcpyAsync per copy on a private stream, CUDA-event timed; `e2e` moves 74 B per "
                    "event per step per rank through the same path"}


def sweep_leg(m, stream, n, rank, world, dev, barrier, args):
    """BASELINE.json configs[4] (C5): 1e7 ... 1e11 generated rays of the C2 set-up in steps of `world` blocks of n rays (the last step
    short), ray indices 64 bit, arrival times continued on the device; detected events counted by a device-resident tally that the
    ranks sum with ONE marxb200_tally_allreduce at the end of each size.  Device timed, max over ranks."""
    import torch
    import torch.distributed as dist
    sizes = [s for s in (10**7, 10**8, 10**9, 10**10, 10**11) if s <= args.sweep_max]
    tally = m.tally_create(("ccd", 10, 0, 10))
    out = []
    for total in sizes:
        tally.reset()
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record(stream)
        first, steps, base = 0, 0, 0.0
        with torch.cuda.stream(stream):
            while first < total:
                cnt = min(n * world, total - first)
                if world == 1:
                    m.trace(first, cnt, base)
                else:
                    m.trace_sharded(first, cnt, base)
                tally.accumulate()
                first += cnt
                steps += 1
                base = -1.0                        # continue the running arrival-time sum on the device
            if world > 1:
                tally.allreduce()
        t1.record(stream)
        barrier()
        ms = t0.elapsed_time(t1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        events = int(tally.read().sum())
        out.append({"rays": total, "steps": steps, "seconds": ms * 1e-3, "rays_per_s": total / (ms * 1e-3), "events": events,
                    "ray_index_bits": 64 if total > (1 << 32) else 32})
    return {"sizes": out, "note": "device resident; the reference counts rays in an int (marx.c:86-87,789) and tags in 32 bits "
                                  "(marx.h:98): beyond 2^32 rays only the 64-bit ray index keys the draws"}


def config_leg(pack, n, stream, dev, args, steps=10):
    """device-resident rays/s of another BASELINE.json configuration at the bench's step size (rank 0, N = 1)"""
    import torch
    import marx_b200
    with marx_b200.MarxB200(pack, device=dev.index, seed=args.seed, max_photons=n, stream=stream.cuda_stream) as c:
        with torch.cuda.stream(stream):
            for k in range(3):
                c.trace(k * n, n)
            torch.cuda.synchronize(dev)
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            t0.record(stream)
            for k in range(steps):
                c.trace((3 + k) * n, n)
            t1.record(stream)
            torch.cuda.synchronize(dev)
        ms = t0.elapsed_time(t1)
        return {"calpack": pack, "rays_per_s": float(n) * steps / (ms * 1e-3), "ms_per_step": ms / steps, "steps": steps,
                "stage_counts_last_step": c.stage_counts()}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.first = index, None, [], 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        self.first = len(self.lines)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)                      # make sure at least one sample falls after the timed region started
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        time.sleep(0.15)
        for ln in self.lines[self.first:]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2]); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "samples": len(sm),
                "power_w_max": max(power) if power else None, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def reference_full_program(per_proc, nprocs, seed0=1000):
    """the reference's own multi-core mode end to end: N independent stock `marx` processes, each writing its output
    directory (tmpfs), time to the last exit, then `marxcat` merging the N directories in time order (marxcat.c:491-535), timed
    separately.  -> dict"""
    import shutil
    ref, _, par, data = ref_paths()
    marx, marxcat = os.path.join(ref, "marx"), os.path.join(ref, "marxcat")
    if not (os.path.exists(marx) and os.path.exists(marxcat)):
        return {"unavailable": "oracle/_ref/marx or marxcat not built"}
    base = scratch_dir("refrun")
    os.makedirs(base)
    env = dict(os.environ, MARX_DATA_DIR=data, USER=os.environ.get("USER", "marx"))
    try:
        t0 = time.time()
        procs = [subprocess.Popen([marx, "@@" + par, "OutputDir=" + os.path.join(base, "run_%d" % k), "NumRays=%d" % per_proc,
                                   "RandomSeed=%d" % (seed0 + k)] + REF_ARGS, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env)
                 for k in range(nprocs)]
        rcs = [p.wait() for p in procs]
        wall = time.time() - t0
        if any(rcs):
            return {"unavailable": "a stock marx process failed"}
        t0 = time.time()
        p = subprocess.run([marxcat] + [os.path.join(base, "run_%d" % k) for k in range(nprocs)] + [os.path.join(base, "merged")],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
        t_cat = time.time() - t0
        rows = None
        f = os.path.join(base, "merged", "pha.dat")
        if p.returncode == 0 and os.path.exists(f):
            rows = (os.path.getsize(f) - 32) // 2
        return {"processes": nprocs, "rays_per_process": per_proc, "wall_s_to_last_exit": wall, "rays_per_s": per_proc * nprocs / wall,
                "marxcat_s": t_cat, "marxcat_ok": p.returncode == 0, "merged_events": rows,
                "rays_per_s_including_marxcat": per_proc * nprocs / (wall + t_cat),
                "note": "whole stock program incl. calibration-file reading and event-file output to tmpfs (default OutputVectors)"}
    finally:
        shutil.rmtree(base, ignore_errors=True)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    per_proc = args.ref_rays_per_proc
    # warm-up (page cache, CPU clocks), then K bounded samples on all host cores
    for _ in range(args.warmup):
        run_reference_sample(max(per_proc // 10, 20000), cores)
    rays_tot, t_tot = 0, 0.0
    for s in range(args.steps):
        rays, secs, wall = run_reference_sample(per_proc, cores, seed0=1 + s * cores)
        rays_tot += rays
        t_tot += secs
    value = rays_tot / t_tot
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": "%d rays per process x %d processes per step" % (per_proc, cores)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "cpu": cpu_model(),
                         "sample": "%d steps x %d processes x %d rays, stock MARX stages + stock RNG, trace loop only "
                                   "(oracle/_ref/marx_trace_bench, gcc -O2)" % (args.steps, cores, per_proc)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    # BASELINE.md 3 extras, reported beside the headline: the -O3 / AVX2+FMA build of the same harness, one core alone, and the whole
    # stock program (output files + marxcat merge)
    extras = {}
    try:
        rays, secs, _ = run_reference_sample(per_proc, cores, seed0=501, exe_name="marx_trace_bench_o3")
        extras["O3_x86_64_v3_all_cores"] = {"value": rays / secs, "unit": UNIT, "cores": cores,
                                            "sample": "same harness, reference compiled gcc -O3 -march=x86-64-v3, %d processes x %d rays" % (cores, per_proc)}
    except Exception as e:  # noqa: BLE001
        extras["O3_x86_64_v3_all_cores"] = {"unavailable": str(e)[:160]}
    try:
        rays, secs, _ = run_reference_sample(per_proc, 1, seed0=601)
        extras["O2_one_core"] = {"value": rays / secs, "unit": UNIT, "cores": 1, "sample": "%d rays, gcc -O2" % rays}
    except Exception as e:  # noqa: BLE001
        extras["O2_one_core"] = {"unavailable": str(e)[:160]}
    try:
        extras["full_program_and_marxcat"] = reference_full_program(max(per_proc // 2, 100000), cores)
    except Exception as e:  # noqa: BLE001
        extras["full_program_and_marxcat"] = {"unavailable": str(e)[:160]}
    line["reference_extras"] = extras
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def phase(msg):
    """progress marker on stderr (MARXB200_BENCH_TRACE=1): where a multi-rank run stands if it ever stalls"""
    if os.environ.get("MARXB200_BENCH_TRACE"):
        sys.stderr.write("[bench rank %s %.1fs] %s\n" % (os.environ.get("RANK", "0"), time.time() - phase.t0, msg))
        sys.stderr.flush()


phase.t0 = time.time()


def cuda_arm(args):
    import faulthandler
    import numpy as np
    if os.environ.get("MARXB200_BENCH_HANG_S"):
        faulthandler.dump_traceback_later(float(os.environ["MARXB200_BENCH_HANG_S"]), exit=True)
    import torch
    import torch.distributed as dist
    import marx_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)      # before any pinned allocation: first touch decides the node
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n = args.rays_per_step
    assert n % 65536 == 0, "rays per step must be a multiple of 65536 (super-tile aligned time sums)"
    stream = torch.cuda.Stream(device=dev)
    m = marx_b200.MarxB200(args.calpack, device=local_rank, seed=args.seed, max_photons=n, stream=stream.cuda_stream)

    # e2e leg: the event list of every step lands in pinned host memory as the column set the reference writes for this
    # configuration by default (marx.par OutputVectors "ETXYZ123DxyMPOabcdSrB" ANDed with the photon history,
    # marxio.c:409-414): 17 float32 + 2 int16 + 2 int8 columns = 74 B per event, in the reference's file encoding
    # (marxb200_egress_begin_packed/_end_packed: converted on the device, copied on a private stream while the next
    # batch is traced).  Two pinned buffers: the host side is double-buffered as well.
    from marx_b200 import HISTORY
    e2e_mask = sum(HISTORY[k] for k in ("ENERGY", "TIME", "X_VECTOR", "P_VECTOR", "DET_NUM", "DET_PIXEL", "MIRROR_SHELL",
                                        "PULSEHEIGHT", "ORDER", "PI", "SKY_DITHER", "DET_DITHER"))
    bytes_per_event = 17 * 4 + 2 * 2 + 2 * 1
    cap = n // 8
    pinned = [torch.empty(cap * (bytes_per_event + 2) + 4096, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    pinned_np = [p.numpy() for p in pinned]

    # multi-GPU: the exchanges run inside the library (comm.cu): NCCL communicator from an id rank 0 creates and the launcher's
    # process group hands over; marxb200_trace_sharded all-gathers the time bases on the device; the event lists are merged
    # on rank 0 over NVLink (marxb200_merge_events_begin/_end) while the next step is traced
    if world > 1:
        from marx_b200.api import comm_unique_id, COMM_ID_BYTES
        idt = torch.zeros(COMM_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, src=0)
        phase("id broadcast done")
        m.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
        phase("comm_init done")
    merge_log = []

    def one_step(step):
        with torch.cuda.stream(stream):
            if world == 1:
                m.trace(step * n, n)           # fused source+HRMA-A, HRMA B1 | B2C1 | C2, grating, detector, order restore
            else:
                m.trace_sharded(step * n * world, n * world)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed(K, e2e, merge=False):
        """K steps.  e2e: every step's event list lands in this rank's pinned host memory (pipelined packed egress).  merge (N>1):
        every step's event lists are merged on rank 0's HBM in arrival order; the transfers of step s run while step s+1 is traced."""
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        launches0 = m.launch_count()
        t0.record(stream)
        n_events = 0
        for s in range(K):
            one_step(timed.step)
            timed.step += 1
            if e2e:
                # pipelined egress: the D2H copy of batch s-1 overlaps the kernels of batch s
                if s > 0:
                    n_events += len(m.egress_end_packed(pinned_np[(s - 1) & 1])["energy.dat"])
                m.egress_begin_packed(e2e_mask, 0.0, cap)
            if merge:
                with torch.cuda.stream(stream):
                    if s > 0:
                        merge_log.append(m.merge_events_end())
                    m.merge_events_begin(e2e_mask, 0.0, cap, 0)
        if e2e:
            n_events += len(m.egress_end_packed(pinned_np[(K - 1) & 1])["energy.dat"])
        if merge:
            merge_log.append(m.merge_events_end())
        t1.record(stream)
        barrier()
        ms = t0.elapsed_time(t1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, m.launch_count() - launches0, n_events
    timed.step = 0

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(1.0)                       # let nvidia-smi come up; it samples every 100 ms from then on
    for _ in range(max(args.warmup, 3)):
        one_step(timed.step); timed.step += 1
    phase("warm-up steps launched")
    if world > 1:
        timed(2, e2e=False, merge=True)       # sets up the merge buffers (collective: IPC mapping of rank 0's buffer)
        merge_log.clear()
    if rank == 0:
        sampler.mark()                        # only samples taken from here on count
    # `value`: at N > 1 the NVLink merge of every step's event lists is inside the timed region
    phase("merge set up")
    ms, launches, _ = timed(args.steps, e2e=False, merge=(world > 1))
    phase("value timed")
    clocks = sampler.stop() if rank == 0 else None
    merge_steps = list(merge_log)
    ms_nomerge = timed(args.steps, e2e=False)[0] if world > 1 else ms
    counts = m.stage_counts()
    # per-kernel durations: the same K steps once more with the library's CUDA-event marks switched on (events
    # recorded on the launching stream after every kernel; the marks cost ~1 % so `value` is timed without them)
    m.set_profiling(True)
    ms_prof, _, _ = timed(args.steps, e2e=False)
    kms = m.kernel_ms()
    m.set_profiling(False)
    # device time per STEP of each kernel class (a class may hold several launches per step: the time scan is 3 kernels, the
    # ACIS detector stage 2, the order restoration 5)
    per = {k: (v[0] / args.steps) for k, v in kms.items()}
    # the mirror stage behind the fused entry runs as B1 | B2+C1 | C2 (cut behind the reflectivity tests) or, with
    # MARXB200_K1_SPLIT=0, as B | C
    k1_four = kms["k1_hrma<B1>"][1] > 0
    k1_names = ["k1_hrma<B1>", "k1_hrma<B2C1>", "k1_hrma<C2>"] if k1_four else ["k1_hrma<1>", "k1_hrma<2>"]
    stage_ms = ([per["k0_time_sums"] + per["k0_time_scan"], per["k01_source_hrma"]] + [per[k] for k in k1_names]
                + [per["k2_grating"], per["k3_detect"], per["order_restore"]])
    # e2e leg
    phase("profiled run done")
    timed(2, e2e=True)
    ms_e2e, _, n_events = timed(args.steps, e2e=True)
    phase("e2e timed")

    # host-copy ceiling of this box (collective): 128 MB pinned D2H copies, every rank alone in turn, then all ranks at once --
    # the bound of `e2e` when 8 ranks land 89 MB per step each in host memory
    d2h = None
    if not args.no_probe:
        try:
            d2h = d2h_probe(m, rank, world, dev, barrier)
        except Exception as e:  # noqa: BLE001
            d2h = {"unavailable": str(e)[:200]}

    phase("d2h probe done")
    # C5: throughput sweep 1e7 ... 1e11 generated rays (collective)
    sweep = None
    if not args.no_sweep:
        sweep = sweep_leg(m, stream, n, rank, world, dev, barrier, args)

    # Level-1 leg (marx2fits' per-event transforms on the device-resident list): the events of the last batch,
    # measured on its own after the timed regions above -- it is not part of `value` / `e2e`
    phase("sweep done")
    one_step(timed.step); timed.step += 1     # a full batch again (the sweep ends on a short one); collective at N > 1
    level1 = None
    if rank == 0 and not args.no_level1:
        try:
            level1 = level1_leg(m, stream, args)
        except Exception as e:  # noqa: BLE001
            level1 = {"unavailable": str(e)[:200]}

    # pile-up leg, likewise on its own after the timed regions
    pileup = None
    if rank == 0 and not args.no_pileup:
        try:
            pileup = pileup_leg(m, stream, args)
        except Exception as e:  # noqa: BLE001
            pileup = {"unavailable": str(e)[:200]}

    # the drop-in driver, run the way a MARX user runs `marx` (rank 0, N = 1): its own process and context
    driver = None
    if rank == 0 and world == 1 and not args.no_driver:
        try:
            driver = driver_leg(args)
        except Exception as e:  # noqa: BLE001
            driver = {"unavailable": str(e)[:200]}

    ic = None
    try:
        ic = [int(v) for v in m.internal_counts()]
    except Exception:
        pass
    fp64_peak, fp64_src = None, None
    if rank == 0:
        try:
            fp64_peak = float(m.measure_fp64_peak())
            fp64_src = "measured in this run (marxb200_measure_fp64_peak: DFMA chains, best of 5)"
        except Exception:  # noqa: BLE001
            fp64_peak, fp64_src = 37.0, "datasheet (measurement failed)"
    comm_info = m.comm_info() if world > 1 else None
    push_rates = None
    if world > 1:
        mine = [x for x in merge_steps if x["copy_ms"] > 0 and x["nvlink_bytes"] > 0]
        rate = (sum(x["nvlink_bytes"] for x in mine) / (sum(x["copy_ms"] for x in mine) * 1e-3) / 1e9) if (mine and rank != 0) else None
        push_rates = [None] * world
        dist.all_gather_object(push_rates, rate)
    phase("legs done")
    if world > 1:
        barrier()                             # every rank leaves the communicator together
    m.close()
    phase("context closed")

    # the other BASELINE.json configurations (C1, C3, C4), device resident, same step size: driver-visible numbers
    configs = None
    if rank == 0 and world == 1 and not args.no_configs:
        configs = {}
        for name, pack in (("C1 point 1.5 keV HRMA+ACIS-S no grating no dither", "c1_acis_s"),
                           ("C3 LETG+HRC-S 0.1-2 keV dither", "c3_letg_hrc_s"),
                           ("C4 BETA source 10' off axis ACIS-I dither", "c4_beta_acis_i")):
            try:
                configs[name] = config_leg(pack, n, stream, dev, args)
            except Exception as e:  # noqa: BLE001
                configs[name] = {"unavailable": str(e)[:200]}

    total_rays = float(n) * world * args.steps
    value = total_rays / (ms * 1e-3)
    e2e_value = total_rays / (ms_e2e * 1e-3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    if "fp64_tflops" in peaks:
        fp64_peak, fp64_src = float(peaks["fp64_tflops"]), "measured (MEASURED_PEAKS.json)"

    # per-kernel roofline table: input rays of each kernel from the device counters of the last step, algorithmic bytes and
    # contract FP64 flop-equivalents per input ray, executed FP64 flops and DRAM bytes per input
    # ray from the committed ncu capture (profiles/r02_kernel_counters.json, tools/ncu_counters.py)
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "r02_kernel_counters.json")))
    except Exception:
        pass
    n_in = {"K0": n, "K01": n, "B1": ic[4] if ic else 0, "B2C1": ic[5] if ic else 0, "C2": ic[7] if ic else 0,
            "K2": ic[1] if ic else 0, "K3": ic[2] if ic else 0, "ORDER": ic[3] if ic else 0}
    if not k1_four and ic:
        n_in["B"], n_in["C"] = ic[4], ic[5]
    rows = ([("K0", "K0 time pre-pass (k0_time_sums + scan)", 0.0, 60.0), ("K01", "K0+K1a fused (k01_source_hrma)", 56.0, 530.0 + 280.0)]
            + ([("B1", "K1 B1 (k1_hrma<3>)", 116.0, 530.0), ("B2C1", "K1 B2+C1 (k1_hrma<4>)", 166.0, 440.0 + 530.0 * 0.96), ("C2", "K1 C2 (k1_hrma<5>)", 166.0, 243.0)]
               if k1_four else [("B", "K1b (k1_hrma<1>)", 116.0, 970.0), ("C", "K1c (k1_hrma<2>)", 166.0, 773.0)])
            + [("K2", "K2 (k2_select + k2_grating<1>)", 114.0, 700.0), ("K3", "K3 (k3_acis<.,1> + k3_acis<.,2>)", 147.0, 850.0),
               ("ORDER", "order restore (5 kernels)", 252.0, 0.0)])
    kernels = {}
    for k, (key, name, bytes_per, flopeq_per) in enumerate(rows):
        t = stage_ms[k] * 1e-3
        ent = {"ms": stage_ms[k], "share": stage_ms[k] / sum(stage_ms), "input_rays": n_in.get(key, 0),
               "algorithmic_bytes_per_input_ray": bytes_per, "contract_flopeq_per_input_ray": flopeq_per}
        if t > 0 and n_in.get(key):
            ent["hbm_frac"] = bytes_per * n_in[key] / t / 1e9 / hbm_peak
            ent["fp64_frac_contract"] = flopeq_per * n_in[key] / t / 1e12 / fp64_peak
            pk = prof.get(key)
            if pk:
                ent["executed_fp64_flop_per_input_ray"] = pk["fp64_flop_per_input_ray"]
                ent["fp64_frac_executed"] = pk["fp64_flop_per_input_ray"] * n_in[key] / t / 1e12 / fp64_peak
                ent["dram_bytes_per_input_ray_ncu"] = pk["dram_bytes_per_input_ray"]
                ent["traffic_over_algorithmic"] = pk["dram_bytes_per_input_ray"] / bytes_per if bytes_per else None
        kernels[name] = ent
    # the dominant kernel by time and its binding roof: every kernel of this path sits above the FP64 / HBM ridge, so
    # the roof is the FP64 pipe; `achieved` counts the FP64 flops the kernel EXECUTES (2 per DFMA, 1 per DMUL / DADD, from the ncu
    # counters of the committed capture) per launch over its CUDA-event time
    top_key, top_name = max(((r[0], r[1]) for r in rows), key=lambda kn: kernels[kn[1]]["ms"])
    top = kernels[top_name]
    flop_per = top.get("executed_fp64_flop_per_input_ray", top["contract_flopeq_per_input_ray"])
    achieved = flop_per * top["input_rays"] / (top["ms"] * 1e-3) / 1e12
    tp = prof.get(top_key)
    # SURVEY 8d contract figures for the whole staged path: 179 B and 1.6e3 FP64 flop-equivalents per generated ray
    path_hbm_gbs = 179.0 * n / (sum(stage_ms) * 1e-3) / 1e9
    path_tflopeq = 1600.0 * n / (sum(stage_ms) * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_step_per_gpu": n, "calpack": args.calpack, "seed": args.seed,
                   "l2": "inputs (1.7 GB photon SoA per batch) exceed the 126 MB L2; no flush needed",
                   "stage_counts_last_step_rank0": counts, "host_affinity_rank0": numa},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 24,
                "d2h_bytes_per_step": int(bytes_per_event * n_events / args.steps) + 8,
                "note": "C ABI marxb200_trace(_sharded) + marxb200_egress_begin_packed/_end_packed: every step's event list (the 21 "
                        "columns the reference writes for this configuration, in its float32/int16/int8 file encoding, 74 B per "
                        "event) is copied to this rank's pinned host memory inside the timed region, overlapped with the next "
                        "batch; the only per-step host input of this path is the batch descriptor (first ray, count, time base)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "fp64", "kernel": top_name, "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": achieved / fp64_peak, "peak_source": fp64_src,
                     "achieved_contract": top["contract_flopeq_per_input_ray"] * top["input_rays"] / (top["ms"] * 1e-3) / 1e12,
                     "frac_contract": top.get("fp64_frac_contract"),
                     "achieved_note": "`achieved` counts the FP64 flops the kernel executes (ncu counters of the committed capture: 2 per "
                                      "DFMA, 1 per DMUL / DADD); `achieved_contract` the SURVEY 8d flop-equivalents per ray (a sqrt, a "
                                      "division or a sin counted at its published cost)",
                     "traffic": (tp["dram_bytes_per_input_ray"] * top["input_rays"]) if tp else None,
                     "flop_per_input_ray": flop_per, "input_rays_per_launch": top["input_rays"], "avg_launch_ms": top["ms"],
                     "hbm": {"achieved_gbs": top["algorithmic_bytes_per_input_ray"] * top["input_rays"] / (top["ms"] * 1e-3) / 1e9,
                             "peak_gbs": hbm_peak, "frac": top.get("hbm_frac"), "peak_source": peak_src},
                     "note": "dominant kernel by time, on the roof that binds it: arithmetic intensity of every kernel of this path is "
                             "above the FP64/HBM ridge of 5.2 flop/B (SURVEY 8d); both fractions per kernel in `kernels`",
                     "whole_path": {"hbm_gbs_at_179B_per_ray": path_hbm_gbs, "hbm_frac": path_hbm_gbs / hbm_peak,
                                    "fp64_tflopeq_at_1600_per_ray": path_tflopeq, "fp64_peak_tflops": fp64_peak,
                                    "fp64_frac": path_tflopeq / fp64_peak},
                     "profiled_ms_per_step": ms_prof / args.steps,
                     "kernels": kernels},
    }
    if world > 1:
        recv = [x for x in merge_steps if x["nvlink_bytes"] > 0 and x["transfer_ms"] > 0]
        tot_b = sum(x["nvlink_bytes"] for x in recv)
        tot_ms = sum(x["transfer_ms"] for x in recv)
        src = [g for g in push_rates[1:] if g]
        line["merge"] = {"what": "every step's per-GPU event lists concatenated in arrival order in rank 0's HBM (marxb200_merge_events_begin/"
                                 "_end, reference analogue marxcat.c:505-535), inside the timed region of `value`, overlapped with the next step",
                         "transport": comm_info["merge_transport"], "nccl_version": comm_info["nccl_version"],
                         "steps": len(merge_steps), "rows_per_step": (sum(x["n_rows"] for x in merge_steps) / max(len(merge_steps), 1)),
                         "nvlink_bytes_per_step_into_rank0": tot_b / max(len(recv), 1),
                         "transfer_ms_per_step": tot_ms / max(len(recv), 1),
                         "achieved_nvlink_gbs_into_rank0": (tot_b / (tot_ms * 1e-3) / 1e9) if tot_ms > 0 else None,
                         "transfer_ms_note": "rank 0's merge-stream time from the start of its own copy to the end of the closing barrier: "
                                             "includes waiting for the slowest source rank",
                         "source_push_gbs_per_rank": src,
                         "source_push_gbs_sum": sum(src) if src else None,
                         "source_push_note": "each source rank's bytes / the CUDA-event time of its own copy-engine writes into rank 0's "
                                             "buffer (21 columns, one cudaMemcpyAsync each); measured peer-copy reference 770 GB/s per direction",
                         "value_without_merge": total_rays / (ms_nomerge * 1e-3),
                         "time_base_exchange": "ncclAllGather of the pre-pass's per-65536-ray sums inside marxb200_trace_sharded, added in "
                                               "global ray order on every GPU (no host round trip)"}
    if d2h is not None:
        line["e2e"]["d2h_ceiling"] = d2h
        # the ranks advance in lockstep (one time-base exchange per step), so a step cannot end before the SLOWEST rank's copy has:
        # bytes per step per rank / that rank's rate with all ranks copying
        slowest = min(d2h["concurrent_gbs_per_rank"])
        if slowest > 0:
            floor_ms = line["e2e"]["d2h_bytes_per_step"] / (slowest * 1e9) * 1e3
            e2e_ms = (total_rays / args.steps) / line["e2e"]["value"] * 1e3
            line["e2e"]["d2h_floor"] = {"ms_per_step": floor_ms, "e2e_ms_per_step": e2e_ms, "device_ms_per_step": line["ms_per_step"],
                                        "bound": "host link" if floor_ms > line["ms_per_step"] else "device",
                                        "e2e_over_floor": e2e_ms / max(floor_ms, line["ms_per_step"]),
                                        "note": "floor = this step's D2H bytes of one rank / the slowest rank's rate with all ranks copying "
                                                "at once (probe above); e2e_over_floor = measured e2e step / max(floor, device step)"}
    if sweep is not None:
        line["sweep"] = sweep
    if configs is not None:
        line["configs"] = configs
    if driver is not None:
        line["driver"] = driver
    if level1 is not None:
        line["level1"] = level1
    if pileup is not None:
        line["pileup"] = pileup
    # CPU baseline beside it (N=1 only): the compiled reference on ONE core, bounded sample
    if world == 1 and not args.no_cpu_baseline:
        try:
            rays, secs, _ = run_reference_sample(args.cpu_baseline_rays, 1)
            line["cpu_baseline"] = {"value": rays / secs, "unit": UNIT, "cores": 1, "kind": "reference", "cpu": cpu_model(),
                                    "sample": "%d rays, stock MARX 5.5.3 stages + stock RNG, trace loop only "
                                              "(oracle/_ref/marx_trace_bench, gcc -O2), %.1f s" % (rays, secs)}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "unavailable: %s" % e}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--rays-per-step", type=int, default=1 << 24)
    ap.add_argument("--calpack", default="c2_hetg_acis_s")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cpu-baseline-rays", type=int, default=6000000)
    ap.add_argument("--ref-rays-per-proc", type=int, default=1000000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-level1", action="store_true")
    ap.add_argument("--no-pileup", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-probe", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--no-driver", action="store_true")
    ap.add_argument("--sweep-max", type=float, default=1e11)
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        cuda_arm(args)


if __name__ == "__main__":
    main()
