# final refresh of the headline lines + launch lists (no --set full captures; those are in the gpu_round.sh ncu pass)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
for prec in fp16x3 bf16 bf16x3 fp16; do
  timeout 600 python bench.py --steps 30 --warmup 5 --precision $prec $( [ $prec != fp16x3 ] && echo --no-cpu-baseline ) > gpurun_out/bench_$prec.log 2>&1
done
timeout 900 python bench.py --mode train --steps 20 --warmup 3 > gpurun_out/bench_train.log 2>&1
timeout 900 python bench.py --mode train --steps 20 --warmup 3 --no-graph > gpurun_out/bench_train_eager.log 2>&1
timeout 300 python bench.py --workload clevr567 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/wl_clevr_fp16x3.log 2>&1
timeout 300 python bench.py --workload clevr567 --steps 30 --warmup 5 --no-cpu-baseline --precision bf16 > gpurun_out/wl_clevr_bf16.log 2>&1
timeout 300 python bench.py --mode train --workload room_diverse_fine --steps 20 --warmup 3 > gpurun_out/wl_fine.log 2>&1
timeout 300 python scripts/train_breakdown.py > gpurun_out/train_breakdown.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
UORF_PROFILE_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 600 python bench.py --mode train --steps 2 --warmup 1 --no-graph > gpurun_out/plain4.log 2>&1 &&
UORF_PROFILE_RANGE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_train.csv python bench.py --mode train --steps 2 --warmup 1 --no-graph > gpurun_out/ncu_launch_train.log 2>&1
for f in gpurun_out/bench_*.log gpurun_out/wl_*.log; do echo $f; python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d.get("roofline",{}).get("kernel_ms"), d["clocks"])
except Exception as e:
    print("ERR", e, open(sys.argv[1]).read()[-600:])
PY
done
