timeout 900 python -m pytest tests/test_gpu_kmeans_sets.py tests/test_gpu_numerical_domain.py tests/test_gpu_direct_path.py tests/test_gpu_edge_cases.py tests/test_gpu_initialisers_predict.py -q -m gpu > gpurun_out/pytest_sets_r02l.log 2>&1; tail -15 gpurun_out/pytest_sets_r02l.log
timeout 900 python -m pytest tests/test_gpu_cabi_parity.py tests/test_gpu_oracle_large.py tests/test_gpu_cppyml.py -q -m gpu -k "kmeans or KMeans or km" > gpurun_out/pytest_km_r02l.log 2>&1; tail -8 gpurun_out/pytest_km_r02l.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-responsibilities > gpurun_out/bench_c3_g1_r02l.json 2> gpurun_out/bench_c3_g1_r02l.err; tail -c 600 gpurun_out/bench_c3_g1_r02l.err
timeout 300 python bench.py --workload c5 --points 12500000 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_c5_r02l.json 2> gpurun_out/bench_c5_r02l.err
MLB200_KM_STATS=owner timeout 300 python bench.py --workload c5 --points 12500000 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_c5_owner_r02l.json 2> gpurun_out/bench_c5_owner_r02l.err
python - <<'PY'
import json
for f in ("bench_c3_g1_r02l", "bench_c5_r02l", "bench_c5_owner_r02l"):
    try:
        l = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, l["value"], l["ms_per_step"], l["roofline"]["frac"], l["roofline"]["kernel_ms_avg"], l["roofline"].get("timed_steps_on_direct_kernels"), l.get("e2e", {}).get("value"))
    except Exception as exc:
        print(f, "failed", exc)
PY
for v in "" build/variants/libmlb200_lse2.so build/variants/libmlb200_lse2_fmax.so; do
  echo "== lib ${v:-in-tree}"
  MLB200_LIB=$v timeout 300 python tools/quick_bench.py em:10000000:8:16 em:12500000:16:32 em:10000000:8:8 em:10000000:4:8 2>&1 | tee -a gpurun_out/lse_variants_r02l.jsonl
done
timeout 600 python tools/bm_kmeans.py > gpurun_out/bm_kmeans_r02l.jsonl 2> gpurun_out/bm_kmeans_r02l.err; cat gpurun_out/bm_kmeans_r02l.jsonl; tail -3 gpurun_out/bm_kmeans_r02l.err
