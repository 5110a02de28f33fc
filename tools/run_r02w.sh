# one more pass of the whole GPU suite on a fresh box (reliability: every test under the 300 s safety net), slowest tests listed
timeout 1200 python -m pytest tests/ -q -m gpu --durations=12 > gpurun_out/pytest_gpu_r02w.log 2>&1; tail -20 gpurun_out/pytest_gpu_r02w.log | cut -c1-200
