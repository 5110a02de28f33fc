timeout 600 python -m pytest tests/test_gpu_kmeans_sets.py tests/test_gpu_numerical_domain.py tests/test_gpu_direct_path.py -q -m gpu -x > gpurun_out/pytest_sets_r02l.log 2>&1; tail -15 gpurun_out/pytest_sets_r02l.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-responsibilities > gpurun_out/bench_c3_g1_r02l.json 2> gpurun_out/bench_c3_g1_r02l.err; tail -c 600 gpurun_out/bench_c3_g1_r02l.err
timeout 300 python bench.py --workload c5 --points 12500000 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_c5_r02l.json 2> gpurun_out/bench_c5_r02l.err
python - <<'PY'
import json
for f in ("bench_c3_g1_r02l", "bench_c5_r02l"):
    l = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, l["value"], l["ms_per_step"], l["roofline"]["frac"], l["roofline"]["kernel_ms_avg"], l["roofline"].get("timed_steps_on_direct_kernels"), l.get("e2e", {}).get("value"))
PY
