# per-tensor gradient errors of the decoder backward against the oracle, for the default library and the variants: VARIANTS="a"
mkdir -p gpurun_out
for v in default ${VARIANTS}; do
  if [ $v = default ]; then unset UORF_B200_LIB; else export UORF_B200_LIB=$PWD/uorf_b200/_build/variants/libuorf_b200_$v.so; fi
  echo "== $v"
  timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_kernels.py -m gpu -q -s -k "backward_values_at_training_size or decoder_grads_golden or decoder_grads_vs_oracle" 2>&1 | grep -E "decoder backward|passed|failed|\{" | cut -c1-1200
done
