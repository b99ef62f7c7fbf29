# round-2 multi-GPU visit (gpurun --gpus N): the driver's torchrun invocation of bench.py at N ranks
N=${1:-2}
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
tail -c 7000 gpurun_out/bench_n$N.log; tail -15 gpurun_out/bench_n$N.err
if [ -n "$MULTI_TESTS" ]; then
  timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "tensor_device" > gpurun_out/pytest_n$N.log 2>&1; tail -3 gpurun_out/pytest_n$N.log
fi
