# ncu launch list of the timed steps of the fine-stage train bench
mkdir -p gpurun_out
T="python bench.py --workload room_diverse_fine --steps 2 --warmup 1 --no-graph --no-cpu-baseline"
timeout 600 $T > gpurun_out/plain_f.log 2>&1 &&
UORF_PROFILE_RANGE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches_fine.csv $T > gpurun_out/ncu_launch_fine.log 2>&1
wc -l gpurun_out/launches_fine.csv; tail -c 300 gpurun_out/plain_f.log
