# interleaved A/B of library variants on the TRAINING step (decoder backward kernel): VARIANTS="a b" ROUNDS=2 [VTESTS="..."]
mkdir -p gpurun_out
if [ -n "$VTESTS" ]; then
  timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_round2.py -m gpu -q -x --timeout 400 -k "$VTESTS" > gpurun_out/ab5_tests.log 2>&1
  tail -5 gpurun_out/ab5_tests.log
fi
for r in $(seq 1 ${ROUNDS:-2}); do
for v in default ${VARIANTS}; do
  if [ $v = default ]; then unset UORF_B200_LIB; else export UORF_B200_LIB=$PWD/uorf_b200/_build/variants/libuorf_b200_$v.so; fi
  timeout 900 python bench.py --mode train --steps ${STEPS:-15} --warmup 4 --no-cpu-baseline > gpurun_out/ab5_${v}.log 2>&1
  python - gpurun_out/ab5_${v}.log $v <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    t=d.get("train") or d
    r=t.get("roofline") or {}
    print(f"{sys.argv[2]:14s} bwd_ms {r.get('kernel_ms',0):.3f} frac {r.get('frac',0):.4f} fwd_ms {r.get('decoder_fwd_kernel_ms',0):.3f} step_ms {t['ms_per_step']:.3f} sm_mhz {t['clocks']['sm_mhz']}",
          "fine", round((d.get("train_fine") or {}).get("ms_per_step", 0), 3))
except Exception as e:
    print("ERR", e); print(open(sys.argv[1]).read()[-1500:])
PY
done
done
