# --set full capture of the encoder launches (forward + gradient kernels) of one training step; plain run of the same command first
mkdir -p gpurun_out
T="python bench.py --mode train --steps 2 --warmup 1 --no-graph --no-cpu-baseline"
timeout 600 $T > gpurun_out/plain_e.log 2>&1 &&
UORF_PROFILE_RANGE=1 timeout 1200 ncu --profile-from-start off --set full --clock-control none -k regex:encoder_ -c 28 -o gpurun_out/prof_encoder -f $T > gpurun_out/ncu_full_enc.log 2>&1
ls -la gpurun_out/prof_encoder.ncu-rep
