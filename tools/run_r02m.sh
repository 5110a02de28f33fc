timeout 120 python tools/debug_kms.py 20000 16 48 4 > gpurun_out/debug_kms_a.log 2>&1; tail -30 gpurun_out/debug_kms_a.log
timeout 120 python tools/debug_kms.py 7000 64 40 2 > gpurun_out/debug_kms_b.log 2>&1; tail -16 gpurun_out/debug_kms_b.log
timeout 200 compute-sanitizer --tool memcheck python tools/debug_kms.py 3000 16 48 4 > gpurun_out/debug_kms_san.log 2>&1; grep -E "ERROR SUMMARY|Invalid|at .*kernel" gpurun_out/debug_kms_san.log | head -20
timeout 150 python -m pytest tests/test_gpu_numerical_domain.py -x -q -m gpu --timeout 60 > gpurun_out/pytest_domain_r02m.log 2>&1; tail -30 gpurun_out/pytest_domain_r02m.log
