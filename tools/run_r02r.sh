for v in "" build/variants/libmlb200_nw6c.so build/variants/libmlb200_nw6a.so; do
  echo "== lib ${v:-in-tree}"
  MLB200_LIB=$v timeout 300 python tools/quick_bench.py em:10000000:8:16 em:10000000:8:8 em:10000000:4:8 em:10000000:4:16 2>&1 | tee -a gpurun_out/nw_variants_r02r.jsonl | cut -c1-150
done
