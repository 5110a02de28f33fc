# Final validation of round 2 on one B200: smoke, whole GPU suite, the driver's bench command and its reference arm, a sustained
# run, c4 on one GPU, a launch list of the default workload's kernels.
set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02v.log 2>&1; tail -2 gpurun_out/smoke_r02v.log
timeout 1500 python -m pytest tests/ -q -m gpu > gpurun_out/pytest_gpu_r02v.log 2>&1; tail -4 gpurun_out/pytest_gpu_r02v.log | cut -c1-200
timeout 800 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_c3_g1_r02v.json 2> gpurun_out/bench_c3_g1_r02v.err; tail -c 400 gpurun_out/bench_c3_g1_r02v.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref_c3_r02v.json 2> gpurun_out/bench_ref_c3_r02v.err; tail -c 300 gpurun_out/bench_ref_c3_r02v.err
timeout 300 python bench.py --gpus 1 --steps 200 --warmup 5 --no-e2e --no-cpu > gpurun_out/bench_c3_g1_sustained_r02v.json 2> gpurun_out/bench_c3_g1_sustained_r02v.err
timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c4_g1_r02v.json 2> gpurun_out/bench_c4_g1_r02v.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c3_r02v.csv python bench.py --points 12500000 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/launches_c3_r02v.log 2>&1
python - <<'PY'
import json
for f in ("bench_c3_g1_r02v", "bench_ref_c3_r02v", "bench_c3_g1_sustained_r02v", "bench_c4_g1_r02v"):
    try:
        l = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        r = l.get("roofline", {})
        print(f, "value", l.get("value"), "ms", l.get("ms_per_step"), "frac", r.get("frac"), "kernel", r.get("kernel_ms_avg"), "direct", r.get("timed_steps_on_direct_kernels"),
              "e2e", l.get("e2e", {}).get("value"), "cpu", l.get("cpu_baseline", {}).get("value"), "clocks", l.get("clocks", {}))
    except Exception as exc:
        print(f, "failed", exc)
PY
