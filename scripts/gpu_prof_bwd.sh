# --set full capture of one decoder_bwd_kernel launch of the training step (after a plain run of the same command)
mkdir -p gpurun_out
T="python bench.py --mode train --steps 2 --warmup 1 --no-graph --no-cpu-baseline"
timeout 600 $T > gpurun_out/plain_t2.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:decoder_bwd_kernel -s 2 -c 1 -o gpurun_out/prof_decoder_bwd -f $T > gpurun_out/ncu_full_bwd.log 2>&1
ls -la gpurun_out/prof_decoder_bwd.ncu-rep
