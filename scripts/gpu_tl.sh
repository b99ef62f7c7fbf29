mkdir -p gpurun_out
for v in timeline timeline_pre; do
  UORF_B200_LIB=$PWD/uorf_b200/_build/variants/libuorf_b200_$v.so timeout 300 python scripts/timeline.py fp16x3 > gpurun_out/tl_$v.log 2>&1; cat gpurun_out/tl_$v.log | tail -22
done
UORF_B200_LIB=$PWD/uorf_b200/_build/variants/libuorf_b200_preload.so timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 300 -k "decoder_golden or decoder_vs_oracle or full_size_render" 2>&1 | tail -3
UORF_B200_LIB=$PWD/uorf_b200/_build/variants/libuorf_b200_preload.so timeout 300 python bench.py --mode render --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('preload(fixed) decoder ms', d['roofline']['kernel_ms'])"
