#!/usr/bin/env python
"""bench.py -- headline benchmark of the PPM-PA hot path (BASELINE.json).

A "step" is ONE whole progressive-photon-mapping pass of the north-star job, BASELINE configs[4]
(ex-glassbox at 1920x1080, 1 M emitted photons, use-classic on, r0 = 0.1): trace photons -> build map ->
eye paths -> direct light -> gather -> combine + accumulate.  The job has 1000 passes whose radius shrinks
along util/iterator.rb:34-38; the K x N timed passes are STRIDED over that schedule (global pass
g = (step * N + rank) * (1000 // (K * N))), so every run covers r = 0.1 -> 0.019 and every rank gets the same
mix of early (large-radius) and late (small-radius) passes.  `value` = radiance-gathered pixels / s with
everything resident on the device; `e2e` = the same through the public C ABI with host buffers (scene/camera
structs in, render, pass image read back to pinned host memory) per step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload config5|config2]

N > 1 is launched by torch.distributed.run (one rank per GPU).  Passes are independent: each rank renders its
own passes with its own Philox streams, and the accumulated images are combined by ONE NCCL sum-reduce per frame
through the C ABI (ppm_accum_reduce, inside the timed region).  --impl reference times the CPU oracle
(restatement of the reference; the Rust original cannot be built here) on the host cores.
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
XRES, YRES = 1920, 1080
NPHOTON = 1_000_000
R0 = 0.1
UC = True
JOB_PASSES = 1000  # the north-star job: 1000 passes x 1 M photons
WORKLOAD = ("configs[4] (north star): ex-glassbox.scene at 1920x1080, 1M emitted photons/pass, use-classic on, filter none, r0=0.1; "
            "the timed passes are strided over the 1000-pass radius schedule of util/iterator.rb:34-38 (r = 0.1 -> 0.019)")


def set_workload(name):
    """configs[4] (default: the north-star job) or configs[1] (1024x1024, the same schedule)."""
    global XRES, YRES, WORKLOAD
    if name == "config2":
        XRES, YRES = 1024, 1024
        WORKLOAD = ("configs[1]: ex-glassbox.scene, 1024x1024, 1M emitted photons/pass, use-classic on, filter none, r0=0.1; "
                    "timed passes strided over the 1000-pass radius schedule of util/iterator.rb:34-38")


def global_pass(step, rank, world, steps_total):
    """Pass id (Philox stream AND radius index) of a rank's step: the K x N passes of a run are spread evenly over the
    1000-pass job, interleaved over the ranks."""
    stride = max(1, JOB_PASSES // max(steps_total * world, 1))
    return (step * world + rank) * stride


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
            "passes_per_step": 1, "parallelism": f"pass-sharded x{n_gpus}; within a GPU passes run round-robin on two lanes (ppm_render_passes), "
                                                 "each pass one CUDA graph launch",
            "radius": "iterator.rb schedule indexed by the GLOBAL pass id; timed passes strided over the 1000-pass job, interleaved over the ranks "
                      "(every rank renders the same mix of radii: per-GPU work is fixed as N grows)",
            "l2": "per-pass working set (~0.6 GB of records, sorted map, node lists, images; regenerated every pass) "
                  "exceeds the 126 MB L2; nothing is reused between timed passes"}


# ---------------------------------------------------------------------------
# CPU side (oracle): cpu_baseline leg and --impl reference
# ---------------------------------------------------------------------------
def cpu_pass_sample(rows, cores, step, steps_total=20):
    """`cores` independent single-threaded oracle passes in parallel (the reference's own parallelism,
    util/iterator.rb NPARA), each: full 1M-photon trace + map build + eye trace of ONE band of `rows` image rows.
    The bands of the concurrent passes are spread evenly over the image height (stratified), and the passes are
    spread over the radius schedule like the GPU arm's.  Returns full-pass-equivalent seconds per pass (EXTRAPOLATED
    from the bands: t_trace + t_build + t_eye_band * yres / rows) and the wall time."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    import ppmpa_b200 as P
    orc = oracle_lib.Oracle()
    sc, cam = load_workload()
    radii = P.radius_schedule(R0, JOB_PASSES)
    ids = [min(global_pass(step, t, cores, steps_total), JOB_PASSES - 1) for t in range(cores)]
    r2 = [float(radii[g]) ** 2 for g in ids]
    row0s = [int(round((t + 0.5) * YRES / cores - rows / 2.0)) for t in range(cores)]
    row0s = [max(0, min(YRES - rows, r)) for r in row0s]
    row1s = [r + rows for r in row0s]
    t0 = time.perf_counter()
    times, stats = orc.render_passes_bands(sc, cam, SEED, ids, NPHOTON, r2, UC, row0s, row1s)
    wall = time.perf_counter() - t0
    equiv = [t[0] + t[1] + t[2] * (YRES / rows) for t in times]
    return sum(equiv) / len(equiv), wall


def cpu_baseline_obj(rows, cores, t_equiv):
    return {"value": cores * XRES * YRES / t_equiv, "unit": UNIT, "cores": cores, "kind": "port", "extrapolated": True,
            "sample": f"EXTRAPOLATED: {cores} concurrent single-threaded oracle passes (C++ restatement, -O2 -ffp-contract=off; hash-grid "
                      f"neighbour search, likely faster than the reference's kd-tree), each: full 1M-photon trace + map build + eye "
                      f"trace/direct light/gather of one band of {rows} of {YRES} image rows, bands stratified over the image height, "
                      f"passes spread over the radius schedule; pass time extrapolated to {YRES} rows"}


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    # bounded sample: every step traces the full 1 M photons (~1 s per core) plus `rows` image rows; the row count
    # shrinks with --steps so that the whole run stays within a few minutes whatever K the driver passes
    rows = env_int("PPM_BENCH_CPU_ROWS", max(4, min(64, 640 // max(args.steps + min(args.warmup, 1), 1))))
    for w in range(min(args.warmup, 1)):
        cpu_pass_sample(rows, cores, w, args.steps)
    eq = []
    for s in range(args.steps):
        t, _ = cpu_pass_sample(rows, cores, s, args.steps)
        eq.append(t)
    t_equiv = sum(eq) / len(eq)
    val = cores * XRES * YRES / t_equiv
    cb = cpu_baseline_obj(rows, cores, t_equiv)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_equiv * 1000.0 / cores, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": base_config(args.gpus),
            "cpu_baseline": cb, "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "CPU oracle (restatement of the Rust reference; cargo/rustc absent so the original cannot be built). EXTRAPOLATED "
                    "from stratified row bands; ms_per_step = full-pass-equivalent seconds per pass / cores"}
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


def pin_to_gpu_numa_node(index):
    """Run this rank's host threads (and first-touch its pinned buffers) on the CPUs next to its GPU: with 8 ranks the
    per-step image read-back (50 MB) otherwise crosses the socket interconnect for half of the GPUs.  Best effort."""
    try:
        import torch
        p = torch.cuda.get_device_properties(index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        cpus = set()
        for part in open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip().split(","):
            if not part:
                continue
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:
        pass
    return None


def run_gpu(args):
    # libraries (NCCL's version banner ...) may write to fd 1: keep stdout for the ONE JSON line
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    import torch.distributed as dist

    import ppmpa_b200 as P

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the engine has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    host_cpus = pin_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = P.Engine(local)
    sc, cam = load_workload()
    eng.set_scene(sc)
    eng.set_camera(cam)
    eng.accum_reset()
    npix = XRES * YRES
    K, W = args.steps, args.warmup
    radii = P.radius_schedule(R0, JOB_PASSES + 1)
    stream = torch.cuda.ExternalStream(eng.stream, device=dev)

    # the frame reduce goes through the C ABI (ppm_accum_reduce over the engine's own NCCL communicator); the only
    # job of torch.distributed here is to hand rank 0's NCCL id to the other ranks, and the barrier / max-over-ranks
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(P.Engine.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, src=0)
        eng.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))

    def gp(step):
        return global_pass(step, rank, world, K)

    def r2_of(g):                                        # beyond the 1000-pass job (K x N > 1000) the radius stays at its last value
        return float(radii[min(g, JOB_PASSES - 1)]) ** 2

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()

    def batch(steps):
        # one ppm_render_passes call = len(steps) whole passes (pass id = global pass: RNG stream and radius index)
        ids = [gp(s_) for s_ in steps]
        stride = (ids[1] - ids[0]) if len(ids) > 1 else 1
        assert all(ids[i + 1] - ids[i] == stride for i in range(len(ids) - 1))
        eng.iterate(SEED, ids[0], len(ids), NPHOTON, [r2_of(g) for g in ids], UC, pass_stride=max(stride, 1))

    # warm-up: W untimed steps taken from the middle and both ends of the schedule (allocations, calibration, graphs, lanes)
    warm = [0, K - 1, K // 2] + list(range(1, max(W - 2, 1)))
    for s_ in warm[:max(W, 3)]:
        batch([max(0, min(K - 1, s_))])
    batch(list(range(min(K, 2))))                        # both lanes once
    if world > 1:
        eng.accum_reduce(root=0)                         # warm the communicator
    eng.accum_reset()

    # ---- device-resident timed region ----------------------------------------------------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    sampler.mark_begin()
    ev0.record(stream)
    batch(list(range(K)))                                # EXACTLY K steps (passes) in the timed region
    phases, counts = eng.last_pass_stats()               # batch totals over the K passes
    if world > 1:
        eng.accum_reduce(root=0)                         # ONE sum-reduce of (3*W*H + 1) doubles per frame (NCCL, C ABI)
    ev1.record(stream)
    sync_all()
    sampler.mark_end()
    clocks = sampler.stop()
    t_ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    t_ms = float(t_ms.item())
    value = world * npix * K / (t_ms / 1000.0)
    n_acc = eng.accum_read()[1] if rank == 0 else 0

    # cross-check of the device stamps: the same K passes in stream mode, k_gather bracketed by CUDA events on its stream
    ev_gather_ms = None
    if rank == 0 and not args.no_crosscheck:
        lanes0 = eng.get_option("lanes")
        eng.set_option("graph", 0)
        eng.set_option("lanes", 1)
        batch(list(range(K)))
        ev_gather_ms = eng.last_pass_stats()[0]["gather_kernel"] / K
        eng.set_option("graph", 1)
        eng.set_option("lanes", lanes0)
    eng.accum_reset()

    # ---- end-to-end through the C ABI with host buffers --------------------------------
    # Every step: ppm_scene_set + ppm_camera_set (host structs through the ABI; uploaded when they differ from what the
    # context holds), one whole pass, and the pass image read back into pinned host memory (what `ppmpa` prints).
    # Passes are independent, so -- like the reference's NPARA = 4 processes (util/iterator.rb:18) -- E2E_LANES engine
    # contexts per GPU are driven by as many host threads, each doing whole steps through the public API.
    import threading
    E2E_LANES = max(1, env_int("PPM_E2E_LANES", max(2, min(4, (os.cpu_count() or 8) // max(world, 1)))))
    engs = [eng] + [P.Engine(local) for _ in range(E2E_LANES - 1)]
    bufs = [torch.empty((npix, 3), dtype=torch.float64).pin_memory().numpy() for _ in range(E2E_LANES)]
    # bytes per step: scene + camera structs handed over through the ABI and the pass table (H2D); the pass image and the
    # pass report (D2H)
    h2d = (C.sizeof(P._capi.Prim) * sc.nprims + C.sizeof(P._capi.Material) * sc.nmats + C.sizeof(P._capi.Light) * sc.nlights
           + C.sizeof(P._capi.Camera) + 12304)
    d2h = npix * 24 + 184

    def e2e_worker(lane, steps):
        e = engs[lane]
        for s_ in steps:
            e.set_scene(sc)
            e.set_camera(cam)
            g = gp(s_)
            e.iteration(SEED, g, NPHOTON, r2_of(g), UC)
            e.pass_image(bufs[lane])

    for lane in range(E2E_LANES):                        # warm the extra contexts (untimed)
        e2e_worker(lane, [0])
    sync_all()
    t0 = time.perf_counter()
    th = [threading.Thread(target=e2e_worker, args=(lane, range(lane, K, E2E_LANES))) for lane in range(E2E_LANES)]
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
        gather_s = phases["gather_kernel"] / 1000.0 / K      # k_gather (+ heavy parts) alone, device stamps in stream order
        alg_bytes = (49.0 * counts["sum_k"] + 72.0 * counts["gather_nodes"]) / K          # per launch
        achieved = alg_bytes / gather_s / 1e9
        traffic = None
        try:
            if XRES == 1920:                             # the ncu capture is of this workload (profiles/gather_traffic.json says which launch)
                traffic = json.load(open(os.path.join(ROOT, "profiles", "gather_traffic.json"))).get("dram_bytes_per_launch")
        except (OSError, ValueError):
            pass
        # what actually limits k_gather: FP64 issue.  Lane-ops counted: 8 per candidate distance test (3 sub, 3 mul, 2 add)
        # + 8 per accepted photon (n.d: 3 mul 2 add; (wt*power)*-cos: 2 mul; accumulate: 1 add); peak = 64 FP64 lanes per SM
        # per clock without FMA contraction (the engine is compiled -fmad=false for bit parity)
        sm_hz = 1e6 * float(clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0)
        sms = torch.cuda.get_device_properties(local).multi_processor_count
        fp64_ops = (8.0 * counts["candidates"] + 8.0 * counts["sum_k"]) / K
        fp64_peak = 64.0 * sms * sm_hz
        build_s = phases["map_build"] / 1000.0 / K
        build_bytes = 114.0 * counts["stored"] / K
        roofline = {"bound": "hbm", "kernel": "k_gather", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": gather_s * 1000.0,
                    "ms_per_launch_cuda_events_stream_mode": ev_gather_ms,
                    "timing": "%globaltimer stamps written by the pass at the first CTA of k_gather and by a 1-thread kernel right "
                              "after its last heavy part, in stream order inside the CUDA graph, live in the timed region; "
                              "cross-checked by CUDA events around the same launches in stream mode (ms_per_launch_cuda_events_stream_mode)",
                    "fp64_issue_frac": fp64_ops / gather_s / fp64_peak,
                    "fp64_lane_ops_per_launch": fp64_ops, "fp64_peak_lane_ops_per_s": fp64_peak,
                    "candidates_tested_per_launch": counts["candidates"] / K,
                    "photons_accepted_per_launch": counts["sum_k"] / K,
                    "queries_per_launch": counts["gather_nodes"] / K,
                    "map_build": {"bound": "hbm", "algorithmic_bytes_per_pass": build_bytes, "ms_per_pass": build_s * 1000.0,
                                  "achieved": build_bytes / build_s / 1e9, "peak": peak, "unit": "GB/s",
                                  "frac": build_bytes / build_s / 1e9 / peak,
                                  "note": "114 B per stored record (56 B record in + 64 B map out, minus padding); the phase is ~20 "
                                          "small dependent kernels over ~0.5 M records: launch/dependency latency bound, not bandwidth"},
                    "note": "frac is LOGICAL bytes (49 B x sum_q K_q + 72 B x N_q, SURVEY 8d) over the HBM peak and can exceed 1: the map is "
                            "L2-resident and every photon is reused by many queries; physical DRAM traffic is `traffic`.  The limiter is "
                            "FP64 issue: see fp64_issue_frac"}
        cpu = None
        if world == 1 and not args.no_cpu:
            rows = env_int("PPM_BENCH_CPU_ROWS", 24)
            cores = os.cpu_count() or 1
            t_equiv, _ = cpu_pass_sample(rows, cores, 0, 1)
            cpu = cpu_baseline_obj(rows, cores, t_equiv)
        passes_per_s = world * K / (t_ms / 1000.0)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": t_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": base_config(world), "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "checksum": checksum, "host_threads_per_gpu": E2E_LANES,
                        "host_cpus_per_rank": len(host_cpus) if host_cpus else None},
                "gpu_launches": counts["launches"], "roofline": roofline, "cpu_baseline": cpu,
                "passes_accumulated_on_rank0_after_reduce": n_acc,
                "passes_retried_after_overflow": counts["retried"],
                "photons_per_sec": world * NPHOTON * K / (t_ms / 1000.0),
                "photon_trace_only_photons_per_sec": NPHOTON / (phases["photon_trace"] / 1000.0 / K),
                "gather_only_queries_per_sec": counts["gather_nodes"] / K / (phases["gather"] / 1000.0 / K),
                "time_to_1000_passes_s": JOB_PASSES / passes_per_s,
                "global_passes_timed": [gp(s_) for s_ in range(K)] if world == 1 else f"(step*{world}+rank)*{max(1, JOB_PASSES // (K * world))}",
                "phases_ms_per_pass": {k: v / K for k, v in phases.items()},
                "per_pass": {k: v / K for k, v in counts.items()}}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        eng.comm_destroy()
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-crosscheck", action="store_true", help="skip the stream-mode CUDA-event cross-check of the k_gather time")
    ap.add_argument("--workload", default="config5", choices=["config2", "config5"],
                    help="config5 = BASELINE configs[4], the north-star job at 1920x1080 (default); config2 = configs[1], 1024x1024")
    args = ap.parse_args()
    set_workload(args.workload)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
