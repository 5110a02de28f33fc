# Two-rank diagnosis (tools/rank_diag.py), with and without torch.distributed in the process.
set -x
DIAG_TORCH=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 tools/rank_diag.py > gpurun_out/rank_diag_torch_r02k.jsonl 2> gpurun_out/rank_diag_torch_r02k.err
DIAG_TORCH=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29722 tools/rank_diag.py > gpurun_out/rank_diag_notorch_r02k.jsonl 2> gpurun_out/rank_diag_notorch_r02k.err
cat gpurun_out/rank_diag_torch_r02k.jsonl; echo ----; cat gpurun_out/rank_diag_notorch_r02k.jsonl; tail -5 gpurun_out/rank_diag_torch_r02k.err
