# one GPU-box visit: parity tests, smoke, bench (all precisions), ncu launch list + one full capture of the decoder
mkdir -p gpurun_out; rm -f gpurun_out/*.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
for grp in probe composite sample_coords slot_attention decoder_golden decoder_vs_oracle eval_driver; do
  echo "=== $grp" >> gpurun_out/tests.log
  timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s --timeout 240 -k "$grp" 2>&1 | tail -45 >> gpurun_out/tests.log
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
for prec in fp16x3 bf16x3 bf16; do
  timeout 600 python bench.py --steps 30 --warmup 5 --precision $prec $( [ $prec != fp16x3 ] && echo --no-cpu-baseline ) > gpurun_out/bench_$prec.log 2>&1
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1
if [ "$1" = "ncu" ]; then
  timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
  timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:decoder_fwd -s 3 -c 1 -o gpurun_out/prof_decoder -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
fi
tail -n 60 gpurun_out/tests.log; cat gpurun_out/smoke.log | tail -5; tail -n 3 gpurun_out/bench_*.log
