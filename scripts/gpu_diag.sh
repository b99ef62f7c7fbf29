mkdir -p gpurun_out
python scripts/diag_grads.py 0 > gpurun_out/diag_grads.log 2>&1; grep -v Warning gpurun_out/diag_grads.log | head -150
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/bench_all.log 2> gpurun_out/bench_all.err
tail -c 9000 gpurun_out/bench_all.log; tail -5 gpurun_out/bench_all.err
