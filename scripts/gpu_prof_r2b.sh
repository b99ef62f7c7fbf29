# round-2 (second session) profile pass: launch lists of the timed steps (render, train) and --set full captures of the decoder
# forward (fp16x3 half-tile kernel), decoder backward and the tensor-core convolution.  Every ncu command follows a plain run of
# the same line (B200_PROFILING.md).
mkdir -p gpurun_out; rm -f gpurun_out/launches*.csv gpurun_out/prof_*.ncu-rep
R="python bench.py --mode render --steps 2 --warmup 1 --no-cpu-baseline"
T="python bench.py --mode train --steps 2 --warmup 1 --no-graph --no-cpu-baseline"
timeout 600 $R > gpurun_out/plain_r.log 2>&1 &&
UORF_PROFILE_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $R > gpurun_out/ncu_launch.log 2>&1
timeout 600 $T > gpurun_out/plain_t.log 2>&1 &&
UORF_PROFILE_RANGE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches_train.csv $T > gpurun_out/ncu_launch_train.log 2>&1
timeout 600 $R > gpurun_out/plain_r2.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:decoder_fwd -s 3 -c 1 -o gpurun_out/prof_decoder -f $R > gpurun_out/ncu_full.log 2>&1
timeout 600 $T > gpurun_out/plain_t2.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:decoder_bwd_kernel -s 2 -c 1 -o gpurun_out/prof_decoder_bwd -f $T > gpurun_out/ncu_full_bwd.log 2>&1
timeout 600 $T > gpurun_out/plain_t3.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv3x3_x3 -s 14 -c 3 -o gpurun_out/prof_conv -f $T > gpurun_out/ncu_full_conv.log 2>&1
wc -l gpurun_out/launches.csv gpurun_out/launches_train.csv; ls -la gpurun_out/*.ncu-rep
