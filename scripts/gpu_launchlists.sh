# ncu launch lists of exactly the timed steps (profiler range) for the render and the train bench
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
UORF_PROFILE_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 600 python bench.py --mode train --steps 2 --warmup 1 --no-graph > gpurun_out/plain4.log 2>&1 &&
UORF_PROFILE_RANGE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_train.csv python bench.py --mode train --steps 2 --warmup 1 --no-graph > gpurun_out/ncu_launch_train.log 2>&1
wc -l gpurun_out/launches.csv gpurun_out/launches_train.csv; grep -c decoder gpurun_out/launches.csv gpurun_out/launches_train.csv
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 240 -k "cuda_graph" 2>&1 | tail -2
