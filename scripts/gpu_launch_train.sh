# ncu launch list of the timed steps of the train bench (profiler range)
mkdir -p gpurun_out
timeout 600 python bench.py --mode train --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/plain4.log 2>&1 &&
UORF_PROFILE_RANGE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches_train.csv python bench.py --mode train --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_launch_train.log 2>&1
wc -l gpurun_out/launches_train.csv
