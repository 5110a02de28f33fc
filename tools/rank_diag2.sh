set -x
run() { tag=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29730 + RANDOM % 50)) bench.py --gpus 2 --points 25000000 --steps 10 --warmup 3 --no-e2e "$@" > gpurun_out/diag2_$tag.json 2> gpurun_out/diag2_$tag.err; python - <<PY
import json
l = json.loads(open("gpurun_out/diag2_$tag.json").read().strip().splitlines()[-1])
print("$tag", l["ms_per_step"], l["roofline"]["kernel_ms_avg"], [(r["kernel_ms_avg"], r["sm_mhz"]) for r in l["per_rank"]])
PY
}
run clocks
run noclocks --no-clocks
run clocks_again
