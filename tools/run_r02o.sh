timeout 900 python -m pytest tests/test_gpu_kmeans_sets.py tests/test_gpu_cabi_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_cppyml.py tests/test_gpu_direct_path.py -q -m gpu -k "kmeans or KMeans or km or sets" > gpurun_out/pytest_km_r02o.log 2>&1; tail -5 gpurun_out/pytest_km_r02o.log | cut -c1-200
timeout 300 python bench.py --workload c5 --points 12500000 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_c5_r02o.json 2> gpurun_out/bench_c5_r02o.err
MLB200_KM_STATS=owner timeout 300 python bench.py --workload c5 --points 12500000 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_c5_owner_r02o.json 2> gpurun_out/bench_c5_owner_r02o.err
timeout 300 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_c2_r02o.json 2> gpurun_out/bench_c2_r02o.err
python - <<'PY'
import json
for f in ("bench_c5_r02o", "bench_c5_owner_r02o", "bench_c2_r02o"):
    try:
        l = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, l["value"], l["ms_per_step"], l["roofline"]["frac"], l["roofline"]["kernel_ms_avg"], l["roofline"].get("timed_steps_on_direct_kernels"), l.get("e2e", {}).get("value"))
    except Exception as exc:
        print(f, "failed", exc)
PY
