# ncu capture of the counting-sort statistics kernel (C5 share on one GPU), raw + source pages
timeout 400 ncu --set full --clock-control none --import-source on -k regex:km_stats_sorted -c 1 -o gpurun_out/ncu_km_stats_sorted_r02p -f python tools/quick_bench.py km:12500000:32:256 > gpurun_out/ncu_km_stats_sorted_r02p.log 2>&1
ncu -i gpurun_out/ncu_km_stats_sorted_r02p.ncu-rep --page raw --csv > gpurun_out/ncu_km_stats_sorted_r02p.csv 2>/dev/null
ncu -i gpurun_out/ncu_km_stats_sorted_r02p.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_km_stats_sorted_r02p_source.csv 2>/dev/null
tail -3 gpurun_out/ncu_km_stats_sorted_r02p.log
