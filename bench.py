#!/usr/bin/env python
"""bench.py -- headline benchmark of the PPM-PA hot path (BASELINE.json).

A "step" is ONE whole progressive-photon-mapping pass of config 2
(ex-glassbox, 1024x1024, 1 M emitted photons, use-classic on, r0 = 0.1 with the
iterator.rb schedule): trace photons -> build map -> eye paths -> direct light
-> gather -> combine + accumulate.  `value` = radiance-gathered pixels / s with
everything resident on the device; `e2e` = the same through the public C ABI
with host buffers (scene/camera set + render + pass image read back) per step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched by torch.distributed.run (one rank per GPU).  Passes are
independent: rank g renders its own passes with its own Philox streams, and the
accumulated images are combined by one NCCL reduce per frame (inside the timed
region).  --impl reference times the CPU oracle (restatement of the reference,
the Rust original cannot be built here) on the host cores.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "radiance-gathered pixels/sec (whole PPM-PA pass: photon trace + map build + eye paths + direct light + gather + accumulate)"
UNIT = "pixels/s"
SEED = 0x5EED0001
XRES = YRES = 1024
NPHOTON = 1_000_000
R0 = 0.1
UC = True
WORKLOAD = "configs[1]: ex-glassbox.scene, 1024x1024, 1M emitted photons/pass, use-classic on, filter none, r0=0.1 (iterator.rb schedule)"


STRONG = False     # config5: a fixed job of passes is shared by the ranks (radius follows the GLOBAL pass index)


def set_workload(name):
    """configs[1] (default, the metric's configuration) or configs[4] (the north-star target: 1920x1080,
    passes sharded over the GPUs, radius schedule indexed by the global pass)."""
    global XRES, YRES, WORKLOAD, STRONG
    if name == "config5":
        XRES, YRES, STRONG = 1920, 1080, True
        WORKLOAD = ("configs[4]: ex-glassbox.scene at 1920x1080, 1M emitted photons/pass, use-classic on, filter none, r0=0.1; "
                    "one job of n_gpus x steps passes sharded round-robin over the GPUs, radius = iterator.rb schedule of the global pass")


def env_int(name, dflt):
    try:
        return int(os.environ.get(name, dflt))
    except ValueError:
        return dflt


def load_workload():
    import ppmpa_b200 as P
    sc = P.read_scene(os.path.join(ROOT, "examples", "ex-glassbox.scene"))
    cam = P.read_camera(os.path.join(ROOT, "examples", "camera0.scr"), xreso=XRES, yreso=YRES, progressive=1,
                        pfilter=P.FILTER_NONE)
    return sc, cam


def base_config(n_gpus):
    return {"workload": WORKLOAD, "pixels_per_pass": XRES * YRES, "photons_per_pass": NPHOTON,
            "passes_per_step": 1, "parallelism": f"pass-sharded x{n_gpus}; within a GPU passes run round-robin on two lanes (ppm_render_passes)",
            "radius": ("iterator.rb schedule indexed by the global pass (one shared job)" if STRONG else
                       "iterator.rb schedule indexed by the per-rank step, so per-GPU work is identical at every N"),
            "l2": "per-pass working set (~280 MB of records, sorted map, node lists, images; regenerated every pass) "
                  "exceeds the 126 MB L2; nothing is reused between timed passes"}


# ---------------------------------------------------------------------------
# CPU side (oracle): cpu_baseline leg and --impl reference
# ---------------------------------------------------------------------------
def cpu_pass_sample(rows, cores, step):
    """`cores` independent single-threaded oracle passes in parallel (the reference's own
    parallelism, util/iterator.rb NPARA), each: full 1M-photon trace + map build + eye trace
    of `rows` image rows.  Returns full-pass-equivalent seconds per pass."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    import ppmpa_b200 as P
    orc = oracle_lib.Oracle()
    sc, cam = load_workload()
    r2 = [float(P.radius_schedule(R0, step + 1)[step]) ** 2] * cores
    row0 = (YRES - rows) // 2
    t0 = time.perf_counter()
    times, stats = orc.render_passes_parallel(sc, cam, SEED, step * cores, cores, NPHOTON, r2, UC, row0, row0 + rows)
    wall = time.perf_counter() - t0
    equiv = [t[0] + t[1] + t[2] * (YRES / rows) for t in times]
    return sum(equiv) / len(equiv), wall


def cpu_baseline_obj(rows, cores, t_equiv):
    return {"value": cores * XRES * YRES / t_equiv, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{cores} concurrent single-threaded oracle passes (C++ restatement, -O2 -ffp-contract=off), each: "
                      f"full 1M-photon trace + map build + eye trace/direct light/gather of {rows} of {YRES} image rows; "
                      f"pass time extrapolated to {YRES} rows"}


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    # bounded sample: every step traces the full 1 M photons (~1 s per core) plus `rows` image rows; the row count
    # shrinks with --steps so that the whole run stays within a few minutes whatever K the driver passes
    rows = env_int("PPM_BENCH_CPU_ROWS", max(4, min(64, 640 // max(args.steps + min(args.warmup, 1), 1))))
    for w in range(min(args.warmup, 1)):
        cpu_pass_sample(rows, cores, w)
    eq = []
    for s in range(args.steps):
        t, _ = cpu_pass_sample(rows, cores, s)
        eq.append(t)
    t_equiv = sum(eq) / len(eq)
    val = cores * XRES * YRES / t_equiv
    cb = cpu_baseline_obj(rows, cores, t_equiv)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_equiv * 1000.0 / cores, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": base_config(args.gpus),
            "cpu_baseline": cb, "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "CPU oracle (restatement of the Rust reference; cargo/rustc absent so the original cannot be built). "
                    "ms_per_step = full-pass-equivalent seconds per pass / cores"}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampling in the background.  Started BEFORE the warm-up (nvidia-smi takes a second to
    start on an 8-GPU box); only samples whose timestamp falls inside the timed window are reported."""
    FIELDS = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20",
                                       "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        rows = []
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(c[2]), float(c[3]), c[6:10]))
            except ValueError:
                continue
        t0, t1 = self.t0 or 0.0, self.t1 or 1e18
        inside = [r for r in rows if t0 - 0.02 <= r[0] <= t1 + 0.02]
        window = "timed region"
        if not inside:                                   # timed region shorter than one sampling period
            inside = [r for r in rows if t0 - 1.0 <= r[0] <= t1 + 1.0]
            window = "timed region +-1 s"
        if inside:
            reasons = set()
            for r in inside:
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            out.update(sm_mhz=statistics.median([r[1] for r in inside]), sm_max_mhz=max(r[2] for r in inside),
                       reasons=sorted(reasons), samples=len(inside), window=window)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


class DevArray:
    """__cuda_array_interface__ view of engine-owned device memory (for torch.as_tensor)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3, "strides": None}


def run_gpu(args):
    # libraries (NCCL's version banner ...) may write to fd 1: keep stdout for the ONE JSON line
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    import torch.distributed as dist

    import ppmpa_b200 as P
    from ppmpa_b200 import parallel

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the engine has no CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = P.Engine(local)
    sc, cam = load_workload()
    eng.set_scene(sc)
    eng.set_camera(cam)
    eng.accum_reset()
    npix = XRES * YRES
    K, W = args.steps, args.warmup
    radii = P.radius_schedule(R0, (W + K + 1) * (world if STRONG else 1))
    ridx = (lambda step: step * world + rank) if STRONG else (lambda step: step)
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local))
    acc_ptr, acc_n = eng.accum_device()
    acc = torch.as_tensor(DevArray(acc_ptr, acc_n), device=torch.device("cuda", local))

    def one_pass(step):
        # pass id (RNG stream) is globally unique; the radius depends on the per-rank step only
        eng.iteration(SEED, step * world + rank, NPHOTON, float(radii[ridx(step)]) ** 2, UC)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()

    def batch(first_step, nsteps):
        # one ppm_render_passes call = nsteps whole passes; pass ids (RNG streams) are globally unique,
        # the radius depends on the per-rank step only
        eng.iterate(SEED, first_step * world + rank, nsteps, NPHOTON, [float(radii[ridx(first_step + i)]) ** 2 for i in range(nsteps)],
                    UC, pass_stride=world)

    batch(0, W)
    with torch.cuda.stream(stream):
        parallel.reduce_accumulators(acc, dst=0)         # warm the communicator (no-op at N=1)
    eng.accum_reset()

    # ---- device-resident timed region ----------------------------------------------------
    phases = {}
    counts = {}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    sampler.mark_begin()
    ev0.record(stream)
    batch(W, K)                                          # EXACTLY K steps (passes) in the timed region
    phases, counts = eng.last_pass_stats()               # batch totals over the K passes
    with torch.cuda.stream(stream):
        parallel.reduce_accumulators(acc, dst=0)         # ONE sum-reduce of (3*W*H + 1) doubles per frame
    ev1.record(stream)
    sync_all()
    sampler.mark_end()
    clocks = sampler.stop()
    t_ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    t_ms = float(t_ms.item())
    value = world * npix * K / (t_ms / 1000.0)

    # ---- end-to-end through the C ABI with host buffers --------------------------------
    # Every step: ppm_scene_set + ppm_camera_set (host structs -> device), one whole pass, and the pass
    # image read back into pinned host memory (what `ppmpa` prints).  Passes are independent, so -- like the
    # reference's NPARA = 4 processes (util/iterator.rb:18) -- E2E_LANES engine contexts per GPU are driven by as
    # many host threads, each doing whole steps through the public API (PPM_E2E_LANES, default 4).
    import threading
    E2E_LANES = max(1, env_int("PPM_E2E_LANES", max(2, min(4, (os.cpu_count() or 8) // max(world, 1)))))
    engs = [eng] + [P.Engine(local) for _ in range(E2E_LANES - 1)]
    bufs = [torch.empty((npix, 3), dtype=torch.float64).pin_memory().numpy() for _ in range(E2E_LANES)]
    h2d = (C.sizeof(P._capi.Prim) * sc.nprims + C.sizeof(P._capi.Material) * sc.nmats + C.sizeof(P._capi.Light) * sc.nlights
           + C.sizeof(P._capi.Camera))
    d2h = npix * 24

    def e2e_worker(lane, steps):
        e = engs[lane]
        for s in steps:
            e.set_scene(sc)
            e.set_camera(cam)
            e.iteration(SEED, s * world + rank, NPHOTON, float(radii[ridx(s)]) ** 2, UC)
            e.pass_image(bufs[lane])

    for lane in range(E2E_LANES):                        # warm the extra contexts (untimed)
        e2e_worker(lane, [0])
    sync_all()
    t0 = time.perf_counter()
    th = [threading.Thread(target=e2e_worker, args=(lane, range(W + lane, W + K, E2E_LANES))) for lane in range(E2E_LANES)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * npix * K / float(te.item())
    checksum = float(sum(bf.sum() for bf in bufs))
    for e in engs[1:]:
        e.close()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        gather_s = phases["gather_kernel"] / 1000.0 / K      # k_gather alone, CUDA events on the ctx stream
        alg_bytes = (49.0 * counts["sum_k"] + 72.0 * counts["gather_nodes"]) / K          # per launch
        achieved = alg_bytes / gather_s / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "gather_traffic.json"))).get("dram_bytes_per_launch")
        except (OSError, ValueError):
            pass
        roofline = {"bound": "hbm", "kernel": "k_gather", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": gather_s * 1000.0,
                    "note": "algorithmic = 49 B x sum_q K_q + 72 B x N_q (logical gathered bytes, SURVEY 8d); the map is "
                            "L2-sized and every photon is reused by many queries, so DRAM traffic is far below this"}
        cpu = None
        if world == 1 and not args.no_cpu:
            rows = env_int("PPM_BENCH_CPU_ROWS", 64)
            cores = os.cpu_count() or 1
            t_equiv, _ = cpu_pass_sample(rows, cores, W)
            cpu = cpu_baseline_obj(rows, cores, t_equiv)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": t_ms / K, "higher_is_better": True, "scaling": "strong" if STRONG else "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": base_config(world), "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "checksum": checksum, "host_threads_per_gpu": E2E_LANES},
                "gpu_launches": counts["launches"], "roofline": roofline, "cpu_baseline": cpu,
                "photons_per_sec": world * NPHOTON * K / (t_ms / 1000.0),
                "photon_trace_only_photons_per_sec": NPHOTON / (phases["photon_trace"] / 1000.0 / K),
                "gather_only_queries_per_sec": counts["gather_nodes"] / K / (phases["gather"] / 1000.0 / K),
                "time_to_100_passes_s": 100.0 / (world * K / (t_ms / 1000.0)),
                "time_to_1000_passes_s": 1000.0 / (world * K / (t_ms / 1000.0)),
                "phases_ms_per_pass": {k: v / K for k, v in phases.items()},
                "per_pass": {k: v / K for k, v in counts.items()}}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="config2", choices=["config2", "config5"],
                    help="config2 = BASELINE configs[1] (the metric's configuration, default); config5 = configs[4], 1920x1080 sharded job")
    args = ap.parse_args()
    set_workload(args.workload)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
