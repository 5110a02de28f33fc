#!/usr/bin/env python
"""Benchmark of the clustering hot path (BASELINE.json): Gaussian-mixture EM iterations.

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --steps K --warmup W     # the reference algorithm on the host CPU

A "step" is one EM iteration (E-step + M-step + statistics exchange + parameter refresh) over the whole
synthetic data set.  Workload at 1 GPU: BASELINE config 2, full-covariance GMM, N=10M, D=8, K=16; with
--gpus G every rank holds N=10M points (weak scaling, one process per GPU, launched by torchrun).
`--workload c3` selects N=100M/8 per GPU... see WORKLOADS.  metric = point-component pairs per second.

One JSON line on stdout (rank 0).  `value` is timed with CUDA events on the library's stream with the data
resident in HBM; `e2e` is the same metric for a whole fit through the C-ABI from HOST (pinned) buffers,
upload and result download inside the timed region; `roofline` is the fused kernel against the measured
FP64 pipe peak; `cpu_baseline` is the CPU oracle (a port of the reference's single-threaded loops) on a
bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# name -> (kind, points per GPU, D, K).  Per-GPU sizes, so that --gpus 8 is the BASELINE.json configuration:
# c3 = N=100M D=16 K=32 EM, c4 = N=20M D=64 K=64 EM, c5 = N=100M D=32 K=256 K-means (Lloyd iterations).
WORKLOADS = {
    "c2": ("em", 10_000_000, 8, 16),
    "c3": ("em", 12_500_000, 16, 32),
    "c4": ("em", 2_500_000, 64, 64),
    "c5": ("km", 12_500_000, 32, 256),
}
METRIC = "em_point_components_per_second"
METRIC_KM = "kmeans_point_centroids_per_second"
UNIT = "Gpoint*comp/s"
DATA_SEED = 20261018
# FP64 peak of this pool's B200s measured with tools/fp64_peaks.cu (profiles/fp64_peaks_r01.json): a chain of
# mma.sync.m8n8k4.f64 (DMMA), the instruction the kernels issue; DFMA measured 36.5.  MEASURED_PEAKS.json has
# no FP64 figure, so this is the denominator, re-measured live below when the tool binary is present.
FP64_PEAK_TFLOPS_MEASURED = 37.0


def f_em(d):
    """Algorithmic FP64 flops per point-component pair per iteration."""
    return 2 * d * d + 8 * d + 6


def flops_per_iteration(kind, n, d, k):
    """EM: N K F_EM(D).  K-means: N K (3D + 1) for the assignment + N (2D + 1) for the update."""
    return n * k * f_em(d) if kind == "em" else n * k * (3 * d + 1) + n * (2 * d + 1)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample", type=int, default=0, help="points of the CPU sample (0 = default)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md recipe).  The sampler is started
    before the warm-up (nvidia-smi takes a few hundred ms to produce its first line), every line is stamped on arrival,
    and only the lines that fall inside [mark_begin, mark_end] are reported; a region shorter than the sampling period
    reports the sample nearest to it."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines = []
        self.proc = None
        self.index = index
        self.t_begin = self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            deadline = time.time() + 3.0
            while not self.lines and time.time() < deadline:
                time.sleep(0.01)
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        parsed = []
        for stamp, line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                parsed.append((stamp, float(parts[0]), float(parts[1]), float(parts[2]), [n for n, f in zip(names, parts[3:7]) if f.lower().startswith("active")]))
            except ValueError:
                continue
        inside = [r for r in parsed if self.t_begin is not None and self.t_begin <= r[0] <= self.t_end]
        note = "samples inside the timed region"
        if not inside and parsed and self.t_begin is not None:
            mid = 0.5 * (self.t_begin + self.t_end)
            inside = [min(parsed, key=lambda r: abs(r[0] - mid))]
            note = "timed region shorter than the 20 ms sampling period: nearest sample"
        sm = sorted(r[1] for r in inside)
        reasons = sorted({n for r in inside for n in r[4]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max((r[2] for r in inside), default=None),
                "power_w_max": max((r[3] for r in inside), default=None), "samples": len(inside), "reasons": reasons, "note": note,
                "timed_region_ms": None if self.t_begin is None else (self.t_end - self.t_begin) * 1e3}


def cpu_baseline(n_cpu, d, k, steps, warmup, kind="em"):
    """The reference's CPU algorithm on a bounded sample of the same synthetic mixture, single-threaded as the reference
    is: the oracle port (oracle/mlpp_oracle.cpp), timed by its own per-step clock.

    oracle/_ref (the reference's own ML/EM.cpp compiled against the first-party Eigen stand-in) is what pins the port
    (tests/test_oracle_vs_reference.py: bit-identical results), but it is NOT what is timed: the stand-in evaluates every
    Eigen expression eagerly into a heap-allocated temporary (one malloc per point-component pair in EM.cpp:205), which
    real Eigen does not do, so its speed says nothing about the reference.  The port runs the same loops on raw arrays
    and is about 5x faster than the stand-in build: the fair (harder to beat) baseline."""
    import numpy as np
    import oracle
    from tests.datasets import synthetic_gmm
    data, _, _ = synthetic_gmm(n_cpu, d, min(k, 64), seed=DATA_SEED % 1000, spread=10.0)
    init = np.ascontiguousarray(data[:k].T)
    if kind == "em":
        fit = oracle.em_fit(data, k, means_init=oracle.EXPLICIT, explicit_means=init, maximum_steps=warmup + steps,
                            absolute_tolerance=0.0, relative_tolerance=0.0, want_responsibilities=False)
    else:
        fit = oracle.kmeans_fit(data, k, init=oracle.EXPLICIT, explicit_means=init, maximum_steps=warmup + steps, absolute_tolerance=0.0)
    secs = fit.step_seconds[warmup:]
    mean = float(np.mean(secs))
    name = "em_fit" if kind == "em" else "kmeans_fit"
    return {"value": n_cpu * k / mean / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"oracle {name}, N={n_cpu} D={d} K={k}, {len(secs)} timed iterations after {warmup} warm-up, single thread like the reference",
            "ms_per_step": mean * 1e3, "host_cores_available": os.cpu_count(),
            "pinned_by": "oracle/_ref (reference sources + Eigen stand-in): bit-identical, tests/test_oracle_vs_reference.py" if oracle.ref_available() else "reference property tests only"}


_ALL_CORES_SNIPPET = """
import sys, numpy as np
sys.path.insert(0, {root!r})
import oracle
from tests.datasets import synthetic_gmm
n, d, k, steps, warmup, kind, seed = {n}, {d}, {k}, {steps}, {warmup}, {kind!r}, {seed}
data, _, _ = synthetic_gmm(n, d, min(k, 64), seed=seed, spread=10.0)
init = np.ascontiguousarray(data[:k].T)
if kind == "em":
    fit = oracle.em_fit(data, k, means_init=oracle.EXPLICIT, explicit_means=init, maximum_steps=warmup + steps, absolute_tolerance=0.0,
                        relative_tolerance=0.0, want_responsibilities=False)
else:
    fit = oracle.kmeans_fit(data, k, init=oracle.EXPLICIT, explicit_means=init, maximum_steps=warmup + steps, absolute_tolerance=0.0)
print(float(np.mean(fit.step_seconds[warmup:])))
"""


def cpu_all_cores(n_cpu, d, k, steps, warmup, kind):
    """SURVEY.md 8(d), optional figure: the reference algorithm has no threads, so "all host cores" means one
    independent single-threaded fit per core, each on its own sample of n_cpu points (what a user could do by hand with
    N-fold sharded data and no statistics exchange).  Aggregate throughput = cores * n_cpu * K / slowest mean step."""
    import subprocess
    cores = os.cpu_count() or 1
    procs = [subprocess.Popen([sys.executable, "-c", _ALL_CORES_SNIPPET.format(root=ROOT, n=n_cpu, d=d, k=k, steps=steps, warmup=warmup, kind=kind,
                                                                                seed=(DATA_SEED + 1 + i) % 1000)],
                              stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, env=dict(os.environ, OMP_NUM_THREADS="1"))
             for i in range(cores)]
    means = []
    for p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            return None
        means.append(float(out.strip().splitlines()[-1]))
    return {"value": cores * n_cpu * k / max(means) / 1e9, "unit": UNIT, "cores": cores,
            "sample": f"{cores} independent single-threaded oracle fits, N={n_cpu} points each, {steps} timed iterations after {warmup} warm-up"}


def reference_build_timing(n_ref, d, k, kind, steps=3):
    """For the record next to the port's number: the reference's OWN translation units (oracle/_ref: ML/EM.cpp,
    ML/KMeans.cpp, ... compiled against the first-party Eigen stand-in) on a smaller sample of the same mixture.  Whole
    fit / iterations (the reference has no per-step clock).  None when oracle/_ref was not built (no /root/reference)."""
    import time
    import numpy as np
    import oracle
    from tests.datasets import synthetic_gmm
    if not oracle.ref_available():
        return None
    data, _, _ = synthetic_gmm(n_ref, d, min(k, 64), seed=DATA_SEED % 1000, spread=10.0)
    init = np.ascontiguousarray(data[:k].T)
    t0 = time.perf_counter()
    if kind == "em":
        fit = oracle.em_fit(data, k, means_init=oracle.EXPLICIT, explicit_means=init, maximum_steps=steps, absolute_tolerance=0.0,
                            relative_tolerance=0.0, want_responsibilities=False, impl="reference")
    else:
        fit = oracle.kmeans_fit(data, k, init=oracle.EXPLICIT, explicit_means=init, maximum_steps=steps, absolute_tolerance=0.0, impl="reference")
    seconds = time.perf_counter() - t0
    iterations = max(1, fit.iterations)
    return {"value": n_ref * k * iterations / seconds / 1e9, "unit": UNIT, "cores": 1, "kind": "reference",
            "sample": f"oracle/_ref (the reference's own sources + Eigen stand-in), N={n_ref} D={d} K={k}, whole fit of {iterations} iterations including initialisation",
            "note": "the stand-in evaluates Eigen expressions eagerly into heap temporaries, so this is slower than a real-Eigen build; the port above is the baseline"}


CPU_SAMPLE = {"c2": (200_000, 500_000), "c3": (20_000, 50_000), "c4": (4_000, 6_000), "c5": (20_000, 40_000)}   # (--impl reference, cpu_baseline)


def describe(workload, kind, d, k):
    algo = "ml::EM full-covariance GMM" if kind == "em" else "ml::Clustering::KMeans Lloyd iterations"
    return f"{workload}: {algo}, D={d}, K={k}"


def run_reference(args):
    """--impl reference: the reference's algorithm on the host CPU: the oracle port, which oracle/_ref (the reference's
    own sources compiled against the Eigen stand-in) pins bit for bit; see cpu_baseline for why the port is what is
    timed.  ml::EM / KMeans are single-threaded, so is this."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, n_gpu, d, k = WORKLOADS[args.workload]
    n_cpu = args.cpu_sample or CPU_SAMPLE[args.workload][0]
    base = cpu_baseline(n_cpu, d, k, args.steps, args.warmup, kind)
    ref_build = reference_build_timing(max(1000, n_cpu // 10), d, k, kind)
    line = {
        "impl": "reference", "metric": METRIC if kind == "em" else METRIC_KM, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": describe(args.workload, kind, d, k) + f"; CPU sample of N={n_cpu} points (throughput is per point, O(N) per iteration)",
                   "points": n_cpu, "dims": d, "components": k},
        "cpu_baseline": {kk: base[kk] for kk in ("value", "unit", "cores", "kind", "sample", "pinned_by")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if ref_build:
        line["reference_build"] = ref_build
    print(json.dumps(line), flush=True)


class EmRunner:
    """EM through the C-ABI: a step is expectation + maximisation + statistics exchange + parameter refresh."""
    kernel = "fused E+M kernel (em_small_kernel for D <= 8, em_kernel for D = 16) or the split E / M kernels (D > 16 or K > 32)"

    def __init__(self, cabi, np, data, k, init_means):
        self.np, self.k, self.init = np, k, init_means
        self.obj = cabi.Em(data, k)
        cov = self.obj.sample_covariance()
        self.obj.set_params(init_means, np.repeat(cov[None], k, axis=0), np.full(k, 1.0 / k))

    def run(self, steps, want=False):
        return self.obj.run_steps(steps, want_ll=want)

    def step_sync(self):
        return self.obj.step()          # reads the log-likelihood back, as EM::fit's convergence test needs

    def results(self, labels_out=None):
        params = self.obj.get_params()
        _, labels = self.obj.emit(want_responsibilities=False, want_labels=True, labels_out=labels_out)
        return params, labels

    def result_bytes(self, n_local, d):
        return n_local * 4 + (d * self.k + self.k * d * d + self.k) * 8

    def param_bytes(self, d):
        return (d * self.k + self.k * d * d + self.k) * 8 + d * d * 8


class KmRunner:
    """K-means through the C-ABI: a step is assignment_step + update_step (KMeans.cpp:80-109)."""
    kernel = "km_assign_kernel (DMMA filter + exact refinement) + km_stats_kernel"

    def __init__(self, cabi, np, data, k, init_means):
        self.np, self.k = np, k
        self.obj = cabi.Km(data, k)
        self.obj.set_centroids(init_means)
        self.last = None

    def run(self, steps, want=False):
        out = []
        for _ in range(steps):
            inertia, changed = self.obj.assign()   # the host needs `changed` for KMeans.cpp:85, so every step reads it back
            shift = self.obj.update()
            out.append(inertia)
        return self.np.array(out)

    def step_sync(self):
        inertia, _ = self.obj.assign()
        self.obj.update()
        return inertia

    def results(self, labels_out=None):
        return self.obj.get_centroids(), self.obj.get_labels(labels_out)

    def result_bytes(self, n_local, d):
        return n_local * 4 + d * self.k * 8

    def param_bytes(self, d):
        return d * self.k * 8


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from ml_b200 import cabi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    uid = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        box = [cabi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    kind, n_per_gpu, d, k = WORKLOADS[args.workload]
    Runner = EmRunner if kind == "em" else KmRunner
    n_total = n_per_gpu * world
    ctx = cabi.Context.for_rank(local_rank, rank, world, uid)
    data = cabi.Data.generate_gmm(ctx, n_total, d, min(k, 64), seed=DATA_SEED)
    _, n_local, _ = data.shape

    # Initial means / centroids: K data points at fixed global indices (0..K-1, held by rank 0), identical on every rank.
    box = [data.download(0, k) if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    init_means = np.ascontiguousarray(box[0].T)

    run = Runner(cabi, np, data, k, init_means)

    # ---- value: K steps with the data resident in HBM, CUDA events on the library's stream, max over ranks
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    run.run(args.warmup)
    launches0 = run.obj.launch_count
    run.obj.set_kernel_timing(True)
    barrier()
    clocks.mark_begin()
    ctx.timer_start()
    trace = run.run(args.steps, True)
    ms_total = ctx.timer_stop()
    barrier()
    clocks.mark_end()
    kernel_ms, kernel_launches = run.obj.kernel_time_ms()
    run.obj.set_kernel_timing(False)
    launches = run.obj.launch_count - launches0
    clock_info = clocks.stop() if rank == 0 else None
    ms_total = max_over_ranks(ms_total)
    ms_per_step = ms_total / args.steps
    value = n_total * k / (ms_per_step * 1e-3) / 1e9
    kernel_ms_avg = max_over_ranks(kernel_ms / max(1, kernel_launches))

    # ---- e2e: a whole fit through the C-ABI from pinned host memory (upload, init, K iterations with the convergence
    # scalars read back every iteration, parameters and labels downloaded), wall clock, max over ranks
    e2e = None
    if not args.no_e2e:
        host = torch.empty((n_local, d), dtype=torch.float64, pin_memory=True)
        begin, _ = cabi.shard_range(n_total, world, rank)
        host_np = host.numpy()
        host_np[:] = data.download(begin, n_local)
        labels_pinned = torch.empty(n_local, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)   # the result buffer, pinned like the input
        run.obj.close(); data.close()
        run = data = None

        def one_fit():
            barrier()
            t0 = time.perf_counter()
            d2 = cabi.Data.upload(ctx, host_np, n_total=n_total)
            r2 = Runner(cabi, np, d2, k, init_means)
            last = 0.0
            for _ in range(args.steps):
                last = r2.step_sync()
            params, labels = r2.results(labels_pinned)
            ctx.synchronize()
            barrier()
            dt = time.perf_counter() - t0
            bytes_out = r2.result_bytes(n_local, d)
            bytes_par = r2.param_bytes(d)
            r2.obj.close(); d2.close()
            return dt, last, bytes_out, bytes_par

        one_fit()  # warm-up (allocator, page faults of the staging paths)
        dt, last_e2e, bytes_out, bytes_par = one_fit()
        dt = max_over_ranks(dt)
        h2d = n_local * d * 8 + bytes_par
        d2h = bytes_out + args.steps * 16
        e2e = {"value": n_total * k * args.steps / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
               "fit_seconds": dt, "iterations": args.steps,
               "what": "one whole fit through the C-ABI from pinned host memory: upload of the rank's points, initialisation, "
                       f"{args.steps} iterations each reading back the convergence scalars, parameters and N labels downloaded (labels into a pinned buffer); "
                       "bytes are per iteration (totals / iterations)",
               "last_scalar": last_e2e}
        if kind == "em" and args.steps - 1 >= args.warmup:
            # iteration `steps` of the e2e fit is iteration `steps - warmup` of the timed resident run (same data, same start)
            assert abs(last_e2e - float(trace[args.steps - 1 - args.warmup])) <= 1e-12 * abs(last_e2e), "the e2e fit and the resident run disagree"

    if rank == 0:
        flops_per_launch = flops_per_iteration(kind, n_per_gpu, d, k)
        achieved = flops_per_launch / (kernel_ms_avg * 1e-3) / 1e12
        peak = FP64_PEAK_TFLOPS_MEASURED
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak = json.load(open(peaks_file))["hbm_gbs"] if os.path.exists(peaks_file) else 6650.0
        bytes_per_point = 8 * d if kind == "em" else 8 * d + 4
        hbm_gbs = n_per_gpu * bytes_per_point / (kernel_ms_avg * 1e-3) / 1e9
        line = {
            "metric": METRIC if kind == "em" else METRIC_KM, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": describe(args.workload, kind, d, k) + f", N={n_per_gpu} per GPU ({n_total} total), "
                                   "initial means = data points 0..K-1" + (", initial covariances = sample covariance" if kind == "em" else ""),
                       "points_total": n_total, "points_per_gpu": n_per_gpu, "dims": d, "components": k,
                       "parallelism": f"points sharded over {world} GPU(s), one ncclAllGather of the sufficient statistics per iteration",
                       "l2": f"input is {n_per_gpu * d * 8 / 1e6:.0f} MB per GPU, larger than the 126 MB L2; no flush needed between iterations",
                       "iterations_per_second": 1e3 / ms_per_step},
            "roofline": {"bound": "tensor", "pipe": "FP64 tensor pipe (DMMA, mma.sync.m8n8k4.f64); shares the SM's FP64 unit with DFMA",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "peak_source": "measured on this pool with tools/fp64_peaks.cu (profiles/fp64_peaks_r01.json); MEASURED_PEAKS.json has no FP64 figure",
                         "kernel": Runner.kernel, "kernel_ms_avg": kernel_ms_avg, "kernel_launches_timed": kernel_launches,
                         "flops_per_launch": flops_per_launch,
                         "flops_per_point_component": f_em(d) if kind == "em" else 3 * d + 1,
                         "hbm_achieved_gbs": hbm_gbs, "hbm_peak_gbs": hbm_peak, "hbm_frac": hbm_gbs / hbm_peak,
                         "traffic": None},
            "clocks": clock_info,
            "gpu_launches": launches,
            "last_scalar": float(trace[-1]),
        }
        prof = os.path.join(ROOT, "profiles", "traffic_r01.json")
        if os.path.exists(prof):
            line["roofline"]["traffic"] = json.load(open(prof)).get(args.workload)
        if e2e is not None:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu:
            n_cpu = args.cpu_sample or CPU_SAMPLE[args.workload][1]
            line["cpu_baseline"] = cpu_baseline(n_cpu, d, k, 3, 1, kind)
            all_cores = cpu_all_cores(max(1000, n_cpu // 4), d, k, 3, 1, kind)
            if all_cores:
                line["cpu_baseline"]["all_cores"] = all_cores
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
