# A/B of library tuning variants (uorf_b200/build.py --tag): bench in two precisions per variant.  VARIANTS="hint200 hint2000"
mkdir -p gpurun_out
for v in default ${VARIANTS}; do
  if [ $v = default ]; then unset UORF_B200_LIB; else export UORF_B200_LIB=$PWD/uorf_b200/_build/variants/libuorf_b200_$v.so; fi
  [ -n "$VTESTS" ] && timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 240 -k "$VTESTS" 2>&1 | tail -1
  for prec in ${PRECS:-fp16x3 bf16}; do
    timeout 600 python bench.py --steps 30 --warmup 5 --precision $prec --no-cpu-baseline > gpurun_out/var_${v}_$prec.log 2>&1
    python - gpurun_out/var_${v}_$prec.log $v $prec <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], sys.argv[3], "ms", round(d["ms_per_step"],3), "dec_ms", round(d["roofline"]["kernel_ms"],3), "frac", round(d["roofline"]["frac"],4), d["clocks"]["sm_mhz"])
except Exception as e:
    print("ERR", e); print(open(sys.argv[1]).read()[-1500:])
PY
  done
done
