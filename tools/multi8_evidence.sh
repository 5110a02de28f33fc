set -x
nvidia-smi -L | wc -l > gpurun_out/multi8_r02i_gpus.txt
( timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -rs > gpurun_out/pytest_multi_g8_r02i.log 2>&1 ) &
( timeout 900 python -m pytest tests/test_gpu_oracle_large.py -q -m gpu -rs --durations=10 > gpurun_out/pytest_oracle_large_g8_r02i.log 2>&1 ) &
( timeout 900 python -m pytest "tests/test_gpu_full_size.py::test_em_c3_at_full_size_through_moment_identities" -q -m gpu -rs > gpurun_out/pytest_full_size_g8_r02i.log 2>&1 ) &
wait
tail -3 gpurun_out/pytest_multi_g8_r02i.log gpurun_out/pytest_oracle_large_g8_r02i.log gpurun_out/pytest_full_size_g8_r02i.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_c3_g8_r02i.json 2> gpurun_out/bench_c3_g8_r02i.err
tail -c 1500 gpurun_out/bench_c3_g8_r02i.json
