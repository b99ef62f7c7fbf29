# interleaved A/B of library variants (boxes differ in sustained SM clock: compare within one call, several rounds)
# VARIANTS="a b"  ROUNDS=3
mkdir -p gpurun_out
for r in $(seq 1 ${ROUNDS:-3}); do
for v in default ${VARIANTS}; do
  if [ $v = default ]; then unset UORF_B200_LIB; else export UORF_B200_LIB=$PWD/uorf_b200/_build/variants/libuorf_b200_$v.so; fi
  timeout 600 python bench.py --mode render --steps ${STEPS:-30} --warmup 5 --precision ${PREC:-fp16x3} --no-cpu-baseline > gpurun_out/ab4_${v}.log 2>&1
  python - gpurun_out/ab4_${v}.log $v <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print(f"{sys.argv[2]:14s} dec_ms {r.get('kernel_ms',0):.3f} frac {r.get('frac',0):.4f} step_ms {d['ms_per_step']:.3f} sm_mhz {d['clocks']['sm_mhz']}")
except Exception as e:
    print("ERR", e); print(open(sys.argv[1]).read()[-1500:])
PY
done
done
