set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29751 tools/upload_ranks.py > gpurun_out/upload_ranks_g8_r02u.json 2> gpurun_out/upload_ranks_g8_r02u.err; cat gpurun_out/upload_ranks_g8_r02u.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29752 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_c3_g8_r02u.json 2> gpurun_out/bench_c3_g8_r02u.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/bench_c3_g8_r02u.json").read().strip().splitlines()[-1])
print("c3 g8 value", l["value"], "ms", l["ms_per_step"], "e2e", l["e2e"]["value"], l["e2e"]["fit_seconds"])
PY
