# round-2 GPU visit: parity tests, smoke, the driver's bench invocations (product + reference arm).  Outputs -> gpurun_out/
mkdir -p gpurun_out; rm -f gpurun_out/*.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
if [ -z "$SKIP_BENCH" ]; then
  ( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/bench_all.log 2> gpurun_out/bench_all.err
  tail -c 6000 gpurun_out/bench_all.log; tail -5 gpurun_out/bench_all.err
  ( time timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/bench_reference.log 2> gpurun_out/bench_reference.err
  tail -c 1500 gpurun_out/bench_reference.log; tail -4 gpurun_out/bench_reference.err
fi
