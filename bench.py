#!/usr/bin/env python
"""bench.py -- frames/s of the 832x624 composite modulate + demodulate hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # product (CUDA) arm
    python bench.py --impl reference --gpus N ...            # reference C on the host cores

A "step" is one pass of the hot path over one batch: every monitor of the batch gets one
crt_modulate + crt_demodulate pair (= one field = one "frame" of the metric).
Workload (BASELINE.json configs[1]): NTSC, 832x624 BGRA in -> 832x624 BGRA out, interlaced (field
alternates every step), full colour, noise 0, blend 1, scanlines 1 -- the CLI's settings
(crt_main.c:221-255).  Synthetic seeded-random images, one distinct image per monitor.

Timed numbers
  value      frames/s, whole job, images resident in HBM, CUDA events, max over ranks
  e2e        same metric through the crtx_frames_host C-ABI call with pinned HOST buffers: the
             H2D copy of every source image and the D2H copy of every decoded image are inside the
             timed region
  roofline   the line kernel (k_lines = crt_core.c:511-664): algorithmic bytes / its mean CUDA-event
             launch duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference C code (oracle/_ref, else the oracle port), 1 thread, bounded sample
  dropin     informational (SURVEY 8d "drop-in fps"): crt_modulate + crt_demodulate of the reference's own interface on
             host buffers, one struct CRT, synchronous, wall clock; measured in a child process after the timed regions
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

# rank 0 prints exactly one line on stdout: keep NCCL's version banner (NCCL_DEBUG=VERSION) off it
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W_IN, H_IN, W_OUT, H_OUT = 832, 624, 832, 624
VARIANT = "ntsc"
METRIC = "frames/sec (832x624 modulate+demodulate)"


def rank_info():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def demod_bytes(field, blend=1, scanlines=1, outw=W_OUT, outh=H_OUT, bpp=4, lines=240, input_size=238420):
    """Algorithmic bytes of one crt_demodulate: read analog, write inp, blend-read the
    240 computed rows, write computed + duplicated rows (crt_core.c:428-432, 584-608, 662-664)."""
    ratio = (((outh << 16) // lines) + 32768) >> 16
    off = (field & 1) * (ratio // 2)
    rows_written = 0
    rows_computed = 0
    for k in range(lines):
        beg = k * outh // lines + off
        end = (k + 1) * outh // lines + off
        if beg >= outh:
            continue
        end = min(end, outh)
        rows_computed += 1
        rows_written += max(1, end - scanlines - beg)
    return 2 * input_size + bpp * outw * (rows_computed * blend + rows_written)


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.001)

    def finish(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# --------------------------------------------------------------------------------------------
# reference / cpu_baseline arm
# --------------------------------------------------------------------------------------------

def _cpu_engine():
    import support as S
    if S.have_ref(VARIANT):
        return "reference", (lambda: S.RefEngine(VARIANT, W_OUT, H_OUT))
    return "port", (lambda: S.OracleEngine(VARIANT, W_OUT, H_OUT))


_WORKER = {}


def _cpu_init(seed=1):
    """Per-process set-up (outside any timed region): one reference instance, one source image."""
    import support as S
    kind, make = _cpu_engine()
    eng = make()
    eng.set(blend=1, scanlines=1)
    _WORKER["eng"] = eng
    if VARIANT not in ("nes", "nes_p0", "nes_p1"):
        _WORKER["img"] = S.rand_image(W_IN, H_IN, seed=seed + os.getpid() % 97)
    else:
        _WORKER["img"] = S.nes_image(W_IN, H_IN, seed=seed + os.getpid() % 97)
    _WORKER["f"] = 0


def _cpu_worker(fields):
    """Run `fields` modulate+demodulate pairs of the bench workload on this host thread."""
    from ntsc_crt_b200 import layout
    if "eng" not in _WORKER:
        _cpu_init()
    eng, img = _WORKER["eng"], _WORKER["img"]
    t0 = time.perf_counter()
    for _ in range(fields):
        f = _WORKER["f"]
        if VARIANT.startswith("nesrgb"):
            eng.modulate(img, format=layout.PIX_BGRA, dot_crawl_offset=f & 1, hue=0)
        elif not VARIANT.startswith("nes"):
            eng.modulate(img, format=layout.PIX_BGRA, as_color=1, field=f & 1, frame=(f >> 1) & 1)
        else:
            eng.modulate(img, dot_crawl_offset=f & 1, hue=0)
        eng.demodulate(24 if VARIANT == "vhs" else 0)
        _WORKER["f"] = f + 1
    return time.perf_counter() - t0


def cpu_baseline_single(seconds=8.0):
    """Reference C path, ONE thread pinned to one core, on a bounded sample of the same workload:
    20 warm-up fields, then at least 200 fields; crt_modulate and crt_demodulate are also timed separately
    (medians, file I/O does not exist here)."""
    import statistics
    import pkgload
    pkgload.load()
    from ntsc_crt_b200 import layout
    kind, _ = _cpu_engine()
    pinned = None
    try:  # like `taskset -c`: keep the scheduler from migrating the measurement
        allowed = sorted(os.sched_getaffinity(0))
        pinned = allowed[len(allowed) // 2]
        os.sched_setaffinity(0, {pinned})
    except (AttributeError, OSError):
        allowed = None
    try:
        _cpu_worker(20)  # warm-up
        eng, img = _WORKER["eng"], _WORKER["img"]
        t_mod, t_dem = [], []
        fields, spent = 0, 0.0
        while spent < seconds or fields < 200:
            f = _WORKER["f"]
            t0 = time.perf_counter()
            if VARIANT.startswith("nesrgb"):
                eng.modulate(img, format=layout.PIX_BGRA, dot_crawl_offset=f & 1, hue=0)
            elif not VARIANT.startswith("nes"):
                eng.modulate(img, format=layout.PIX_BGRA, as_color=1, field=f & 1, frame=(f >> 1) & 1)
            else:
                eng.modulate(img, dot_crawl_offset=f & 1, hue=0)
            t1 = time.perf_counter()
            eng.demodulate(24 if VARIANT == "vhs" else 0)
            t2 = time.perf_counter()
            _WORKER["f"] = f + 1
            t_mod.append(t1 - t0)
            t_dem.append(t2 - t1)
            spent += t2 - t0
            fields += 1
    finally:
        if allowed is not None:
            os.sched_setaffinity(0, set(allowed))
    return {"value": fields / spent, "unit": "frames/s", "cores": 1, "kind": kind,
            "sample": "%d fields of the bench workload (832x624 NTSC, noise 0, blend 1), 1 thread%s, %.1f s"
                      % (fields, " pinned to cpu %d" % pinned if pinned is not None else "", spent),
            "modulate_ms_median": 1e3 * statistics.median(t_mod), "demodulate_ms_median": 1e3 * statistics.median(t_dem),
            "host_cores": os.cpu_count()}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation on all the host cores it can use
    (one single-threaded instance per core: the library keeps its filter state in file statics,
    crt_core.c:158-164, so it is process- not thread-parallel)."""
    import multiprocessing as mp
    import pkgload
    pkgload.load()
    rank, local_rank, world = rank_info()
    if rank != 0:
        return
    kind, _ = _cpu_engine()
    cores = os.cpu_count() or 1
    fields_per_worker = 16
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init) as pool:
        jobs = [fields_per_worker] * cores
        for _ in range(max(1, args.warmup)):
            pool.map(_cpu_worker, jobs, chunksize=1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_worker, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    frames = args.steps * cores * fields_per_worker
    value = frames / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i32",
        "data": "synthetic",
        "config": {"workload": ("NTSC 832x624 BGRA -> 832x624 BGRA, interlaced, colour, noise 0, blend 1, scanlines 1" if VARIANT == "ntsc"
                                else "%s -> 832x624 BGRA, blend 1, scanlines 1 (informational run of another BASELINE config)" % VARIANT),
                   "batch_per_step": cores * fields_per_worker, "host_processes": cores},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": "%d processes x %d fields per step" % (cores, fields_per_worker)},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# product arm
# --------------------------------------------------------------------------------------------

def dropin_fps(seconds=2.0):
    """SURVEY 8d "drop-in fps": the reference's own seven-function interface on HOST buffers, exactly what the
    unmodified drivers call -- crt_modulate + crt_demodulate on one struct CRT, synchronous, strict coherence (the
    library re-uploads analog[] and the image on every call, crt_dropin.cu).  Wall clock around the calls, like a
    caller sees it.  Informational: it never touches the contract's `value` / `e2e`, and a failure here is reported
    in the key instead of costing the line."""
    try:
        import support as S
        from ntsc_crt_b200 import layout
        eng = S.ProductEngine(VARIANT, W_OUT, H_OUT)
        eng.set(blend=1, scanlines=1)
        nes = VARIANT in ("nes", "nes_p0", "nes_p1")
        img = S.nes_image(W_IN, H_IN, seed=7) if nes else S.rand_image(W_IN, H_IN, seed=7)

        def pair(f):
            if nes:
                eng.modulate(img, dot_crawl_offset=f & 1, hue=0)
            elif VARIANT.startswith("nesrgb"):
                eng.modulate(img, format=layout.PIX_BGRA, dot_crawl_offset=f & 1, hue=0)
            else:
                eng.modulate(img, format=layout.PIX_BGRA, as_color=1, field=f & 1, frame=(f >> 1) & 1)
            eng.demodulate(24 if VARIANT == "vhs" else 0)
        for f in range(4):
            pair(f)
        n, t0 = 0, time.perf_counter()
        while True:
            pair(n)
            n += 1
            dt = time.perf_counter() - t0
            if (dt >= seconds and n >= 8) or n >= 4000:
                break
        return {"value": n / dt, "unit": "frames/s", "pairs": n, "wall_s": dt,
                "api": "crt_modulate + crt_demodulate (crt_core.h:100-139) on host buffers, one struct CRT, synchronous, strict coherence"}
    except Exception as e:  # informational key: never take the bench line down
        return {"value": None, "error": "%s: %s" % (type(e).__name__, e)}


def dropin_isolated(seconds=2.0):
    """dropin_fps in a child process (`bench.py --impl dropin`): the drop-in library abort()s when CUDA fails (its
    signatures have no error channel), and an informational figure must not be able to take this process with it."""
    import subprocess
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "dropin", "--variant", VARIANT,
                            "--dropin-seconds", str(seconds)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=180)
        if r.returncode != 0:
            return {"value": None, "error": "child exit %d: %s" % (r.returncode, r.stderr.strip()[-200:])}
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:
        return {"value": None, "error": "%s: %s" % (type(e).__name__, e)}


def run_product(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import pkgload
    pkgload.load()
    from ntsc_crt_b200 import capi, layout

    rank, local_rank, world = rank_info()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    gen = torch.Generator(device="cpu").manual_seed(1234 + rank)
    # inputs: one distinct image per monitor, BGRA; far larger than L2 in total (see config)
    nes = VARIANT in ("nes", "nes_p0", "nes_p1")
    noise = 24 if VARIANT == "vhs" else 0  # BASELINE configs[4]: VHS at noise 24
    if nes:
        src = torch.randint(0, 512, (B, H_IN, W_IN), dtype=torch.int16, generator=gen).to(dev)
    else:
        src = torch.randint(0, 256, (B, H_IN, W_IN, 4), dtype=torch.uint8, generator=gen).to(dev)
    out = torch.zeros(B, H_OUT, W_OUT, 4, dtype=torch.uint8, device=dev)
    batch = capi.Batch(VARIANT, B)
    batch.set_option("timing", 1)
    for kv in args.set:
        name, val = kv.split("=")
        batch.set_option(name, int(val))
    for i in range(B):
        batch.set_monitor(i, out[i], fmt=layout.PIX_BGRA, noise=noise, blend=1, scanlines=1)
    batch.commit_monitors()
    # two prebuilt source tables, even / odd field (crt_main.c:245-253 toggles field each pass)
    tables = []
    for field in (0, 1):
        t = (capi.Source * B)()
        for i in range(B):
            s = t[i]
            s.data = src[i].data_ptr()
            s.format, s.w, s.h = layout.PIX_BGRA, W_IN, H_IN
            s.raw, s.as_color, s.field, s.frame = 0, 1, field, 0
            s.dot_crawl_offset, s.reinit = field, 0
        tables.append(t)
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream

    # --streams 2: the batch advances as two halves on two streams, the second one modulate behind the
    # first, so that the issue-bound line pass of one half shares the SMs with the latency-bound encoder
    # and sync search of the other (their register / shared-memory footprints fit side by side).
    halves = None
    if args.streams == 2:
        s2 = torch.cuda.Stream(dev)
        h0 = B // 2
        halves = ((0, h0, sp), (h0, B - h0, s2.cuda_stream))
        stagger = torch.cuda.Event()

    def step(k):
        if halves is None:
            batch._check(batch.lib.crtx_modulate(batch._ctx, 0, B, tables[k & 1], sp))
            batch._check(batch.lib.crtx_demodulate(batch._ctx, 0, B, sp))
            return
        t = tables[k & 1]
        (f0, n0, p0), (f1, n1, p1) = halves
        batch._check(batch.lib.crtx_modulate(batch._ctx, f0, n0, C.cast(C.byref(t, f0 * C.sizeof(capi.Source)), C.POINTER(capi.Source)), p0))
        stagger.record(stream)
        batch._check(batch.lib.crtx_demodulate(batch._ctx, f0, n0, p0))
        s2.wait_event(stagger)
        batch._check(batch.lib.crtx_modulate(batch._ctx, f1, n1, C.cast(C.byref(t, f1 * C.sizeof(capi.Source)), C.POINTER(capi.Source)), p1))
        batch._check(batch.lib.crtx_demodulate(batch._ctx, f1, n1, p1))

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    if nes or VARIANT.startswith("nesrgb"):  # first call of a NES stream writes the sync template (crt_nes.c:118-121)
        t0 = (capi.Source * B)()
        C.memmove(t0, tables[0], C.sizeof(t0))
        for i in range(B):
            t0[i].reinit = 1
        batch._check(batch.lib.crtx_modulate(batch._ctx, 0, B, t0, sp))
    for k in range(args.warmup):
        step(k)
    barrier()
    batch.timing()  # drop warm-up timings
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    launches0 = batch.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for k in range(args.steps):
        step(args.warmup + k)
    if halves is not None:
        stream.wait_stream(s2)  # the timed region ends when both halves are done
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.finish()
    launches = batch.launches - launches0
    ktimes = batch.timing()

    # ---------------- end to end through the host-buffer C-ABI call
    Be = min(args.e2e_batch, B)
    nstreams = 4
    per = Be // nstreams
    if nes:
        h_src = torch.randint(0, 512, (Be, H_IN, W_IN), dtype=torch.int16, generator=gen).pin_memory()
    else:
        h_src = torch.randint(0, 256, (Be, H_IN, W_IN, 4), dtype=torch.uint8, generator=gen).pin_memory()
    h_out = torch.zeros(Be, H_OUT, W_OUT, 4, dtype=torch.uint8).pin_memory()
    streams = [torch.cuda.Stream(dev) for _ in range(nstreams)]
    htables = []
    for field in (0, 1):
        t = (capi.Source * Be)()
        for i in range(Be):
            s = t[i]
            s.data = h_src[i].data_ptr()
            s.format, s.w, s.h = layout.PIX_BGRA, W_IN, H_IN
            s.raw, s.as_color, s.field, s.frame = 0, 1, field, 0
        htables.append(t)
    outp = (C.c_void_p * Be)(*[h_out[i].data_ptr() for i in range(Be)])

    def e2e_step(k):
        for q in range(nstreams):
            first = q * per
            batch._check(batch.lib.crtx_frames_host(
                batch._ctx, first, per,
                C.cast(C.byref(htables[k & 1], first * C.sizeof(capi.Source)), C.POINTER(capi.Source)),
                C.cast(C.byref(outp, first * C.sizeof(C.c_void_p)), C.POINTER(C.c_void_p)),
                streams[q].cuda_stream))

    e2e_steps = max(2, min(2 * args.steps, 32))  # long enough that filling / draining the 4-stream pipeline is noise
    for k in range(2):
        e2e_step(k)
    barrier()
    launches_e0 = batch.launches
    t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(e2e_steps):
        e2e_step(k)
    ends = []
    for q in range(nstreams):
        e = torch.cuda.Event(enable_timing=True)
        e.record(streams[q])
        ends.append(e)
    barrier()
    e2e_ms = max(e0.elapsed_time(e) for e in ends)
    e2e_wall = time.perf_counter() - t0
    e2e_ms = max(e2e_ms, 0.0)
    launches_e2e = batch.launches - launches_e0
    batch.timing()
    frames_e2e = per * nstreams * e2e_steps

    # ---------------- optional: all_gather of the decoded frames (the exchange north_star mentions for
    # "a batch of frames presented together"); priced separately, it is NVLink-bound
    gather = None
    if args.allgather and world > 1:
        from ntsc_crt_b200 import sharding
        g_steps = max(2, min(args.steps, 4))
        full = sharding.allgather_frames(out)  # warm-up, also allocates
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for k in range(g_steps):
            step(k)
            full = sharding.allgather_frames(out)
        g1.record(stream)
        barrier()
        g_ms = sharding.max_over_ranks([g0.elapsed_time(g1)], device=dev)[0]
        gather = {"value": world * B * g_steps / (g_ms / 1e3), "unit": "frames/s",
                  "bytes_received_per_rank_per_step": (world - 1) * B * H_OUT * W_OUT * 4, "steps": g_steps}
        del full

    # ---------------- max over ranks
    if world > 1:
        t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])
    value = world * B * args.steps / (ms / 1e3)
    e2e_value = world * frames_e2e / (e2e_ms / 1e3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        lines_ms, lines_n = ktimes["lines"]
        isz = batch.spec.input_size
        per_launch_bytes = B * (demod_bytes(0, input_size=isz) + demod_bytes(1, input_size=isz)) / 2.0
        achieved = (per_launch_bytes / 1e9) / ((lines_ms / max(1, lines_n)) / 1e3) if lines_n else None
        kernel_share = {k: round(v[0] / ms, 4) for k, v in ktimes.items()}
        traffic = None
        try:  # DRAM bytes of the line kernel per launch, from the committed ncu --set full capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj["k_lines_fir_dram_bytes_per_frame" if ("_conv" in VARIANT) else "k_lines_dram_bytes_per_frame"] * B
        except Exception:
            pass
        cpu = cpu_baseline_single() if (world == 1 and not args.no_cpu_baseline) else None
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
            "config": {"workload": ("NTSC 832x624 BGRA -> 832x624 BGRA, interlaced, colour, noise 0, blend 1, scanlines 1 (BASELINE configs[1])"
                                    if VARIANT == "ntsc" else "%s -> 832x624 BGRA, noise %d, blend 1, scanlines 1 (informational run of another BASELINE config)" % (VARIANT, noise)),
                       "batch_per_gpu": B, "global_batch": B * world, "streams": args.streams, "parallelism": "dp%d (frames sharded, no collective)" % world,
                       "l2": "inputs larger than L2: %.0f MB of images + signals touched per step per GPU" % (B * 4.63)},
            "e2e": {"value": e2e_value, "unit": "frames/s",
                    "h2d_bytes_per_step": per * nstreams * W_IN * H_IN * (2 if nes else 4),
                    "d2h_bytes_per_step": per * nstreams * W_OUT * H_OUT * 4,
                    "batch": per * nstreams, "steps": e2e_steps, "api": "crtx_frames_host, pinned host buffers, %d streams" % nstreams,
                    "wall_s": e2e_wall},
            "gpu_launches": int(launches),
            "gpu_launches_e2e": int(launches_e2e),
            # the CLI accumulates 8 modulate + demodulate pairs per interlaced image (crt_main.c:242-255)
            "cli_images_per_s": value / 8.0,
            "roofline": {"bound": "hbm", "kernel": ("k_lines_fir (crt_demodulate line pass of the USE_CONVOLUTION build, crt_core.c:96-147,511-664)"
                                                     if ("_conv" in VARIANT) else "k_lines (crt_demodulate line pass, crt_core.c:511-664)"),
                         "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                         "algorithmic_bytes_per_launch": per_launch_bytes, "launch_ms": lines_ms / max(1, lines_n),
                         "peak_source": peak_src},
            "kernel_ms_per_step": {k: round(v[0] / max(1, args.steps), 4) for k, v in ktimes.items()},
            "kernel_share_of_step": kernel_share,
            "clocks": clocks,
        }
        if world == 1:
            line["dropin"] = dropin_isolated(0.3 if args.no_cpu_baseline else 2.0)
        if cpu:
            line["cpu_baseline"] = cpu
        if gather:
            line["allgather"] = gather
        print(json.dumps(line), flush=True)
    batch.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference", "dropin"],
                    help="dropin: only the informational drop-in figure (used by the product arm in a child process)")
    ap.add_argument("--dropin-seconds", type=float, default=2.0)
    ap.add_argument("--streams", type=int, default=1, choices=[1, 2],
                    help="2: advance the batch as two halves on two CUDA streams, staggered (see step())")
    ap.add_argument("--batch", type=int, default=296,
                    help="monitors (frames per step) per GPU; 296 = 2 resident CTAs x 148 SMs of the line kernel")
    ap.add_argument("--e2e-batch", type=int, default=64)
    ap.add_argument("--set", action="append", default=[], help="library option name=value (A/B testing)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--allgather", action="store_true", help="also time steps that all_gather the decoded frames")
    ap.add_argument("--variant", default="ntsc", choices=["ntsc", "ntsc_conv", "ntsc_conv6", "ntsc_conv5", "ntsc_conv4", "nes", "nes_p0", "nes_p1", "snes", "nesrgb", "nesrgb_p0", "nesrgb_p1", "vhs", "template", "pv1k", "ntsc_bloom"],
                    help="informational runs of the other systems (the contract metric is the default, ntsc)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "product" else args.warmup
    global VARIANT, W_IN, H_IN
    VARIANT = args.variant
    if VARIANT.startswith("nes"):
        W_IN, H_IN = 256, 240  # PPU image (BASELINE configs[2])
    if args.impl == "dropin":
        import pkgload
        pkgload.load()
        print(json.dumps(dropin_fps(args.dropin_seconds)), flush=True)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
