# full GPU suite, then ncu captures of the kernels changed in r02 (C2 fused kernel, C3 fused kernel, counting-sort statistics)
timeout 1500 python -m pytest tests/ -q -m gpu -x > gpurun_out/pytest_gpu_r02q.log 2>&1; tail -5 gpurun_out/pytest_gpu_r02q.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:em_small_kernel -c 1 -s 3 -o gpurun_out/ncu_em_c2_full_r02q -f python tools/quick_bench.py em:10000000:8:16 > gpurun_out/ncu_c2_r02q.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:km_stats_sorted -c 1 -s 2 -o gpurun_out/ncu_km_stats_sorted_r02q -f python tools/quick_bench.py km:12500000:32:256 > gpurun_out/ncu_kms_r02q.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c5_r02q.csv python tools/quick_bench.py km:12500000:32:256 > gpurun_out/launches_c5_r02q.log 2>&1
for f in ncu_em_c2_full_r02q ncu_km_stats_sorted_r02q; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.csv 2>/dev/null
  ncu -i gpurun_out/$f.ncu-rep --page source --csv --print-source sass > gpurun_out/${f}_source.csv 2>/dev/null
done
ls -la gpurun_out/*r02q*
