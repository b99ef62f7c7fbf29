# A/B of decoder-forward variants: parity tests of the default build, then bench --mode render per variant library
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 300 -k "decoder or eval_driver or full_size_render or manip or row_blocks" > gpurun_out/pytest_dec.log 2>&1; tail -4 gpurun_out/pytest_dec.log
for v in default ${VARIANTS:-nopp legacy}; do
  if [ $v = default ]; then unset UORF_B200_LIB; else export UORF_B200_LIB=$PWD/uorf_b200/_build/variants/libuorf_b200_$v.so; fi
  for prec in ${PRECS:-fp16x3}; do
    timeout 300 python bench.py --mode render --steps 20 --warmup 5 --no-cpu-baseline --precision $prec > gpurun_out/ab_${v}_$prec.log 2>&1
    python - gpurun_out/ab_${v}_$prec.log $v $prec <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print(sys.argv[2], sys.argv[3], "ms/step", round(d["ms_per_step"],3), "decoder ms", round(d["roofline"]["kernel_ms"],3), "frac", round(d["roofline"]["frac"],4))
except Exception as e:
    print(sys.argv[2], "ERR", e); print(open(sys.argv[1]).read()[-1500:])
PY
  done
done
