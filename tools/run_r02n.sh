timeout 600 python -m pytest tests/test_gpu_kmeans_sets.py tests/test_gpu_numerical_domain.py tests/test_gpu_direct_path.py tests/test_gpu_edge_cases.py tests/test_gpu_initialisers_predict.py -v -m gpu --timeout 120 --timeout-method=thread > gpurun_out/pytest_sets_r02n.log 2>&1; grep -E "PASSED|FAILED|ERROR|Timeout|passed|failed" gpurun_out/pytest_sets_r02n.log | tail -25 | cut -c1-200
for v in "" build/variants/libmlb200_lse2_fmax.so; do
  echo "== lib ${v:-in-tree}"
  MLB200_LIB=$v timeout 300 python tools/quick_bench.py em:10000000:8:16 em:12500000:16:32 em:10000000:8:8 em:10000000:4:8 2>&1 | tee -a gpurun_out/lse_variants_r02n.jsonl | cut -c1-200
done
timeout 400 python tools/bm_kmeans.py 100 1000 10000 100000 > gpurun_out/bm_kmeans_r02n.jsonl 2> gpurun_out/bm_kmeans_r02n.err; cat gpurun_out/bm_kmeans_r02n.jsonl | cut -c1-400; tail -3 gpurun_out/bm_kmeans_r02n.err
