mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -x --timeout 300 > gpurun_out/pytest_r2.log 2>&1; tail -4 gpurun_out/pytest_r2.log
VARIANTS="${VARIANTS}" PRECS="${PRECS:-fp16x3}" bash scripts/gpu_ab.sh 2>&1 | grep -v passed
