mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for grp in probe composite sample_coords slot_attention "decoder_golden" "decoder_vs_oracle"; do
  echo "=== $grp" >> gpurun_out/t1.log
  timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s --timeout 240 -k "$grp" 2>&1 | tail -60 >> gpurun_out/t1.log
done
python -c "
from uorf_b200 import _lib
print('debug flag', _lib.lib().uorf_debug_flag())" >> gpurun_out/t1.log 2>&1
tail -150 gpurun_out/t1.log
