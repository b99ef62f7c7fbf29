# mid-round GPU visit: full parity suite, bench (fp16x3 with baselines, bf16), HBM-kernel microbench, ncu full capture of the
# decoder kernel in both modes.  Outputs in gpurun_out/.
mkdir -p gpurun_out; rm -f gpurun_out/*.log gpurun_out/*.ncu-rep
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_fp16x3.log 2>&1
timeout 600 python bench.py --steps 30 --warmup 5 --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16.log 2>&1
timeout 300 python scripts/bench_kernels.py 32 > gpurun_out/bench_kernels.json 2> gpurun_out/bench_kernels.log
for prec in fp16x3 bf16; do
  timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --precision $prec > gpurun_out/plain_$prec.log 2>&1 &&
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:decoder_fwd -s 3 -c 1 -o gpurun_out/prof_decoder_$prec -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --precision $prec > gpurun_out/ncu_full_$prec.log 2>&1
done
for f in gpurun_out/bench_*.log; do echo $f; tail -c 3000 $f | tail -1 | cut -c1-3000; done
cat gpurun_out/bench_kernels.json
