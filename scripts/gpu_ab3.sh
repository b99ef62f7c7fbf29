# round-2 A/B of forward-kernel variants: decoder parity tests on the default build, then the render bench per variant.
# VARIANTS="v1 warparrive"  (tags built with `python -m uorf_b200.build --tag=<tag> -D...`);  MODE=render|strong
mkdir -p gpurun_out
if [ -n "$VTESTS" ]; then
  timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_round2.py -m gpu -q -x --timeout 300 -k "$VTESTS" > gpurun_out/ab3_tests.log 2>&1
  tail -5 gpurun_out/ab3_tests.log
fi
for v in default ${VARIANTS}; do
  if [ $v = default ]; then unset UORF_B200_LIB; else export UORF_B200_LIB=$PWD/uorf_b200/_build/variants/libuorf_b200_$v.so; fi
  for prec in ${PRECS:-fp16x3}; do
    timeout 600 python bench.py --mode ${MODE:-render} --steps ${STEPS:-30} --warmup 5 --precision $prec --no-cpu-baseline > gpurun_out/ab3_${v}_$prec.log 2>&1
    python - gpurun_out/ab3_${v}_$prec.log $v $prec <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print(sys.argv[2], sys.argv[3], "ms", round(d["ms_per_step"],3), "dec_ms", round(r.get("kernel_ms",0),3), "frac", round(r.get("frac",0),4), d["clocks"]["sm_mhz"],
          "strong", [(k, round(v["ms_per_step"],2)) for k,v in (d.get("strong") or {}).items() if isinstance(v, dict) and "ms_per_step" in v])
except Exception as e:
    print("ERR", e); print(open(sys.argv[1]).read()[-1500:])
PY
  done
done
