# 8-GPU evidence (r02s): the multi-GPU tests, then bench.py under torchrun as the driver launches it: c3 (north star, N = 100M
# sharded over 8), c5 (K-means N = 100M through cppyml), c4 (N = 20M, D = 64, K = 64).
set -x
nvidia-smi -L | wc -l > gpurun_out/multi8_r02s_gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -rs > gpurun_out/pytest_multi_g8_r02s.log 2>&1
tail -n 3 gpurun_out/pytest_multi_g8_r02s.log
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29800 + RANDOM % 100)) bench.py --gpus 8 "$@" > gpurun_out/bench_${tag}_g8_r02s.json 2> gpurun_out/bench_${tag}_g8_r02s.err; }
run c3 --steps 20 --warmup 5
run c5 --workload c5 --steps 10 --warmup 3
run c4 --workload c4 --steps 10 --warmup 3
python - <<'PY'
import json
for tag in ("c3", "c5", "c4"):
    try:
        l = json.loads(open(f"gpurun_out/bench_{tag}_g8_r02s.json").read().strip().splitlines()[-1])
        print(tag, "value", l["value"], "ms", l["ms_per_step"], "frac", l["roofline"]["frac"], "kernel", l["roofline"]["kernel_ms_avg"], "direct", l["roofline"].get("timed_steps_on_direct_kernels"),
              "e2e", l.get("e2e", {}).get("value"), l.get("e2e", {}).get("fit_seconds"), "clocks", l["clocks"]["sm_mhz"], l["clocks"]["reasons"])
    except Exception as exc:
        print(tag, "failed", exc)
PY
