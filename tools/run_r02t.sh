set -x
timeout 300 python -m pytest tests/test_gpu_copies.py -q -m gpu > gpurun_out/pytest_copies_r02t.log 2>&1; tail -3 gpurun_out/pytest_copies_r02t.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 tools/upload_ranks.py > gpurun_out/upload_ranks_g2_r02t.json 2> gpurun_out/upload_ranks_g2_r02t.err; cat gpurun_out/upload_ranks_g2_r02t.json; tail -3 gpurun_out/upload_ranks_g2_r02t.err
