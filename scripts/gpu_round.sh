# one full GPU-box visit: parity tests, smoke, bench (all precisions + train + reference arm), probe timing,
# and (with "ncu") the ncu launch list + full captures of the decoder kernel.  Outputs land in gpurun_out/.
mkdir -p gpurun_out; rm -f gpurun_out/*.log gpurun_out/*.csv gpurun_out/*.ncu-rep
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 120 python scripts/probe_timing.py > gpurun_out/probe_timing.log 2>&1
for prec in fp16x3 bf16x3 bf16 fp16; do
  timeout 600 python bench.py --steps 30 --warmup 5 --precision $prec $( [ $prec != fp16x3 ] && echo --no-cpu-baseline ) > gpurun_out/bench_$prec.log 2>&1
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1
timeout 900 python bench.py --mode train --steps 20 --warmup 3 > gpurun_out/bench_train.log 2>&1
timeout 900 python bench.py --mode train --steps 20 --warmup 3 --no-graph > gpurun_out/bench_train_eager.log 2>&1
timeout 300 python scripts/bench_kernels.py 32 > gpurun_out/bench_kernels.json 2> gpurun_out/bench_kernels.err
timeout 120 python scripts/bench_encoder.py > gpurun_out/bench_encoder.log 2>&1
timeout 300 python bench.py --workload clevr567 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/wl_clevr_fp16x3.log 2>&1
timeout 300 python bench.py --workload clevr567 --steps 30 --warmup 5 --no-cpu-baseline --precision bf16 > gpurun_out/wl_clevr_bf16.log 2>&1
timeout 300 python bench.py --mode train --workload room_diverse_fine --steps 20 --warmup 3 > gpurun_out/wl_fine.log 2>&1
timeout 300 python scripts/train_breakdown.py > gpurun_out/train_breakdown.log 2>&1
if [ "$1" = "ncu" ]; then
  timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
  UORF_PROFILE_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
  timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:decoder_fwd -s 3 -c 1 -o gpurun_out/prof_decoder -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --precision bf16 > gpurun_out/plain3.log 2>&1 &&
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:decoder_fwd -s 3 -c 1 -o gpurun_out/prof_decoder_bf16 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --precision bf16 > gpurun_out/ncu_full_bf16.log 2>&1
  timeout 600 python bench.py --mode train --steps 2 --warmup 1 --no-graph > gpurun_out/plain4.log 2>&1 &&
  UORF_PROFILE_RANGE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_train.csv python bench.py --mode train --steps 2 --warmup 1 --no-graph > gpurun_out/ncu_launch_train.log 2>&1
  timeout 300 python scripts/bench_kernels.py 32 > gpurun_out/plain5.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:composite_fwd -s 10 -c 1 -o gpurun_out/prof_composite -f python scripts/bench_kernels.py 32 > gpurun_out/ncu_full_composite.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:sample_coords -s 10 -c 1 -o gpurun_out/prof_sample_coords -f python scripts/bench_kernels.py 32 > gpurun_out/ncu_full_coords.log 2>&1
  timeout 100 python scripts/enc_once.py > gpurun_out/plain7.log 2>&1 &&
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:encoder_conv -s 12 -c 6 -o gpurun_out/prof_encoder -f python scripts/enc_once.py > gpurun_out/ncu_full_encoder.log 2>&1
  timeout 600 python bench.py --mode train --steps 2 --warmup 1 --no-graph > gpurun_out/plain6.log 2>&1 &&
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:decoder_bwd_kernel -s 2 -c 1 -o gpurun_out/prof_decoder_bwd -f python bench.py --mode train --steps 2 --warmup 1 --no-graph > gpurun_out/ncu_full_bwd.log 2>&1
fi
for f in gpurun_out/bench_*.log; do echo $f; tail -c 600 $f | tail -1 | cut -c1-400; done
