# ncu capture of the headline kernel (fused E+M, D = 16, K = 32) on the final build, C3's per-GPU share
timeout 300 ncu --set full --clock-control none --import-source on -k regex:em_kernel -c 1 -s 4 -o gpurun_out/ncu_em_c3_full_r02y -f python tools/quick_bench.py em:12500000:16:32 > gpurun_out/ncu_c3_r02y.log 2>&1
ncu -i gpurun_out/ncu_em_c3_full_r02y.ncu-rep --page raw --csv > gpurun_out/ncu_em_c3_full_r02y.csv 2>/dev/null
ncu -i gpurun_out/ncu_em_c3_full_r02y.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_em_c3_full_r02y_source.csv 2>/dev/null
tail -2 gpurun_out/ncu_c3_r02y.log
