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


def run_reference_sample(nrays_per_proc, nprocs, seed0=1):
    """nprocs concurrent single-threaded reference processes (the reference's own multi-core mode:
    N independent `marx` runs); returns (total rays, wall seconds of the trace loops)."""
    ref, exe, par, data = ref_paths()
    if not os.path.exists(exe):
        raise RuntimeError("oracle/_ref/marx_trace_bench not built (run __graft_entry__.build() where /root/reference exists)")
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
        try:
            from tests import level1_lib
            ph = m.download()
            cols = {"time": ph["arrival_time"].astype(np.float32), "xpixel": ph["y_pixel"], "ypixel": ph["z_pixel"], "b_energy": ph["pi"],
                    "pha": ph["pulse_height"], "ccd": ph["ccd_num"],
                    **{key: np.ascontiguousarray(ph["dither"][:, j]) for j, key in enumerate(level1_lib.DITHER_KEYS)}}
            o = level1_lib.Level1Oracle(desc, args.seed)
            t0 = time.time()
            o.transform(cols)
            dt = time.time() - t0
            out["cpu_baseline"] = {"value": events / dt, "unit": "events/s", "cores": 1, "kind": "port",
                                   "sample": "%d events, oracle/level1_oracle.c (gcc -O2), %.2f s" % (events, dt)}
        except Exception as e:  # noqa: BLE001
            out["cpu_baseline"] = {"value": None, "kind": "port", "sample": "unavailable: %s" % str(e)[:120]}
    return out


def pileup_leg(m, stream, args, n=1 << 22, reps=5):
    """marxb200_pileup_run (marxpileup's frame loop) on a synthetic bright-source event list: ~32 events per
    3.241 s exposure frame on a spot of sigma 3 pixels.  `ms` = the eight kernels (CUDA events inside the library), `e2e_ms` = the
    whole call with HOST columns in and out.  Beside it the pinned plain-C oracle (oracle/pileup_oracle.c, one core) on the same list."""
    import numpy as np
    import torch
    r = np.random.default_rng(args.seed)
    alpha, ft, rate = 0.5, 3.241, 10.0
    cols = {"ccd": np.full(n, 7, np.int8), "t": np.cumsum(r.exponential(1.0 / rate, n)).astype(np.float32),
            "x": (512.0 + r.normal(0.0, 3.0, n)).astype(np.float32), "y": (300.0 + r.normal(0.0, 3.0, n)).astype(np.float32),
            "benergy": r.uniform(0.4, 7.0, n).astype(np.float32)}
    for k in ("sky_ra", "sky_dec", "sky_roll", "det_dy", "det_dz", "det_theta"):
        cols[k] = r.normal(0.0, 1e-3, n).astype(np.float32)
    with torch.cuda.stream(stream):
        got, _ = m.pileup(cols, alpha, ft, args.seed)
        ms, wall = [], []
        for _ in range(reps):
            t0 = time.time()
            got, k = m.pileup(cols, alpha, ft, args.seed)
            wall.append((time.time() - t0) * 1e3)
            ms.append(k)
    ms, wall, rows = float(np.median(ms)), float(np.median(wall)), len(got["t"])
    # algorithmic bytes: R 41 per event (ccd 1, pixel/time/energy 16, aspect 24), W 49 per output row
    out = {"events": n, "rows": rows, "ms": ms, "events_per_s": n / (ms * 1e-3), "e2e_ms": wall, "e2e_events_per_s": n / (wall * 1e-3),
           "launches_per_call": 8, "hbm_gbs_algorithmic": (41.0 * n + 49.0 * rows) / (ms * 1e-3) / 1e9,
           "workload": "synthetic list, %d events, %.0f events/s, frame %.3f s, alpha %.1f, one chip" % (n, rate, ft, alpha)}
    if not args.no_cpu_baseline:
        try:
            from tests import pileup_lib
            t0 = time.time()
            ref = pileup_lib.oracle_pileup(cols, ["Alpha=%r" % alpha, "FrameTime=%r" % ft, "FrameTransferTime=0.0"], args.calpack, args.seed)
            dt = time.time() - t0
            out["parity"] = bool(all(got[k].tobytes() == ref[k].tobytes() for k in ref))
            out["cpu_baseline"] = {"value": n / dt, "unit": "events/s", "cores": 1, "kind": "port",
                                   "sample": "%d events, oracle/pileup_oracle.c (gcc -O2), %.2f s" % (n, dt)}
        except Exception as e:  # noqa: BLE001
            out["cpu_baseline"] = {"value": None, "kind": "port", "sample": "unavailable: %s" % str(e)[:120]}
    return out


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
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": "%d steps x %d processes x %d rays, stock MARX stages + stock RNG, trace loop only "
                                   "(oracle/_ref/marx_trace_bench, gcc -O2)" % (args.steps, cores, per_proc)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def cuda_arm(args):
    import numpy as np
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

    bases = {}

    def plan_time_bases(step0, K):
        """multi-GPU: arrival times are one running sum over ALL rays (source.c:326).  Each rank reduces the time
        increments of its K blocks to super-tile sums (marxb200_time_sums), ONE all-gather over NCCL exchanges them,
        and every rank adds them up in global ray order -- exactly the additions a single GPU performs -- to get the
        absolute time base of each of its blocks.  Single GPU: the running sum simply continues on the device."""
        if world == 1:
            for s in range(step0, step0 + K):
                bases[s] = -1.0
            return
        from marx_b200.dist import block_time_bases
        mine = np.stack([m.time_sums((s * world + rank) * n, n) for s in range(step0, step0 + K)])      # [K, n_super]
        t = torch.from_numpy(mine).to(dev)
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        allsums = torch.stack(gathered).cpu().numpy()                                                      # [world, K, n_super]
        for k in range(K):
            b, plan_time_bases.running = block_time_bases([allsums[r, k] for r in range(world)], plan_time_bases.running)
            bases[step0 + k] = b[rank]
    plan_time_bases.running = 0.0

    def time_base_for(step):
        return (step * world + rank) * n, bases[step]

    def one_step(step):
        first, base = time_base_for(step)
        with torch.cuda.stream(stream):
            m.trace(first, n, base)            # fused source+HRMA-A, HRMA-B, HRMA-C, grating, detector, order restore

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed(K, e2e):
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        launches0 = m.launch_count()
        t0.record(stream)
        plan_time_bases(timed.step, K)        # inside the timed region
        n_events = 0
        for s in range(K):
            one_step(timed.step)
            timed.step += 1
            if e2e:
                # pipelined egress: the D2H copy of batch s-1 overlaps the kernels of batch s
                if s > 0:
                    n_events += len(m.egress_end_packed(pinned_np[(s - 1) & 1])["energy.dat"])
                m.egress_begin_packed(e2e_mask, 0.0, cap)
        if e2e:
            n_events += len(m.egress_end_packed(pinned_np[(K - 1) & 1])["energy.dat"])
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
    plan_time_bases(timed.step, max(args.warmup, 3))
    for _ in range(max(args.warmup, 3)):
        one_step(timed.step); timed.step += 1
    if rank == 0:
        sampler.mark()                        # only samples taken from here on count
    ms, launches, _ = timed(args.steps, e2e=False)
    clocks = sampler.stop() if rank == 0 else None
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
    timed(2, e2e=True)
    ms_e2e, _, n_events = timed(args.steps, e2e=True)

    # Level-1 leg (marx2fits' per-event transforms on the device-resident list): the events of the last batch,
    # measured on its own after the timed regions above -- it is not part of `value` / `e2e`
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
    # dominant kernel among those that read their input from HBM: HRMA phase B1 (optical constants, P-conic normal, blur, Fresnel
    # reflectivity test) -- or the whole phase B when the mirror stage runs uncut.  Algorithmic bytes per INPUT ray of that kernel
    # (DESIGN.md section 4): R x,p 48 + energy 8 + shell/state 3 = 59; W per survivor x,p 48 + state 23 = 71 (B), + normal 24 = 95 (B1)
    ic = None
    try:
        ic = [int(v) for v in m.internal_counts()]
    except Exception:
        pass
    roof_kernel = "k1_hrma<3> (B1)" if k1_four else "k1_hrma<1>"
    k1b_in = ic[4] if ic and ic[4] else int(0.476 * n)
    k1b_out = ic[5] if ic and ic[5] else int((0.62 if k1_four else 0.573) * k1b_in)
    k1b_bytes = 59.0 * k1b_in + (95.0 if k1_four else 71.0) * k1b_out
    k1_ms = per[k1_names[0]]
    achieved = k1b_bytes / (k1_ms * 1e-3) / 1e9
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "k1b_traffic.json")))
        ent = tj.get("all_kernels", {}).get("k1_hrma<3>" if k1_four else "k1_hrma<1>")
        traffic = (ent["dram_read_bytes"] + ent["dram_write_bytes"]) if ent else None
    except Exception:
        pass
    # FP64 peak: not in MEASURED_PEAKS.json -> measured live with the library's DFMA-chain microbenchmark
    fp64_src = "measured (MEASURED_PEAKS.json)"
    if "fp64_tflops" in peaks:
        fp64_peak = float(peaks["fp64_tflops"])
    else:
        try:
            fp64_peak = float(m.measure_fp64_peak())
            fp64_src = "measured in this run (marxb200_measure_fp64_peak: DFMA chains, best of 5)"
        except Exception:  # noqa: BLE001
            fp64_peak, fp64_src = 37.0, "datasheet (measurement failed)"
    names = (["K0 time pre-pass (k0_time_sums+k0_time_scan)", "K0+K1a fused (k01_source_hrma)"]
             + (["K1 B1 (k1_hrma<3>)", "K1 B2+C1 (k1_hrma<4>)", "K1 C2 (k1_hrma<5>)"] if k1_four else ["K1b (k1_hrma<1>)", "K1c (k1_hrma<2>)"])
             + ["K2 (k2_grating)", "K3 (k3_acis<.,1> + k3_acis<.,2>)", "order restore (5 kernels)"])
    kernels = {}
    for k, name in enumerate(names):
        kernels[name] = {"ms": stage_ms[k], "share": stage_ms[k] / sum(stage_ms)}
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
                "note": "C ABI marxb200_trace_from + marxb200_egress_begin_packed/_end_packed: every step's event list (the 21 "
                        "columns the reference writes for this configuration, in its float32/int16/int8 file encoding, 74 B per "
                        "event) is copied to pinned host memory inside the timed region, overlapped with the next batch; the "
                        "only per-step host input of this path is the batch descriptor (first ray, count, time base)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": roof_kernel, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": k1b_bytes, "input_rays_per_launch": k1b_in, "output_rays_per_launch": k1b_out,
                     "avg_launch_ms": k1_ms,
                     "note": "every kernel of this path sits above the FP64/HBM ridge (SURVEY 8d): instruction issue, not HBM, "
                             "binds; the HBM fraction is reported because the contract asks for it",
                     "whole_path": {"hbm_gbs_at_179B_per_ray": path_hbm_gbs, "hbm_frac": path_hbm_gbs / hbm_peak,
                                    "fp64_tflopeq_at_1600_per_ray": path_tflopeq, "fp64_peak_tflops": fp64_peak,
                                    "fp64_frac": path_tflopeq / fp64_peak,
                                    "fp64_peak_source": fp64_src},
                     "profiled_ms_per_step": ms_prof / args.steps,
                     "kernels": kernels},
    }
    if level1 is not None:
        line["level1"] = level1
    if pileup is not None:
        line["pileup"] = pileup
    # CPU baseline beside it (N=1 only): the compiled reference on ONE core, bounded sample
    if world == 1 and not args.no_cpu_baseline:
        try:
            rays, secs, _ = run_reference_sample(args.cpu_baseline_rays, 1)
            line["cpu_baseline"] = {"value": rays / secs, "unit": UNIT, "cores": 1, "kind": "reference",
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
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        cuda_arm(args)


if __name__ == "__main__":
    main()
