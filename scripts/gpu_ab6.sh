# interleaved A/B of the training step under environment settings: ENVS="UORF_ENCODER_BACKWARD=cudnn" ROUNDS=2 [VTESTS="..."]
mkdir -p gpurun_out
if [ -n "$VTESTS" ]; then
  timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_round2.py -m gpu -q -x --timeout 400 -k "$VTESTS" > gpurun_out/ab6_tests.log 2>&1
  tail -15 gpurun_out/ab6_tests.log
fi
for r in $(seq 1 ${ROUNDS:-2}); do
for v in default ${ENVS}; do
  if [ $v = default ]; then e=""; else e="$v"; fi
  env $e timeout 900 python bench.py --mode train --steps ${STEPS:-15} --warmup 4 --no-cpu-baseline > gpurun_out/ab6_${v%%=*}.log 2>&1
  python - gpurun_out/ab6_${v%%=*}.log $v <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    t=d.get("train") or d
    r=t.get("roofline") or {}
    print(f"{sys.argv[2]:34s} bwd_ms {r.get('kernel_ms',0):.3f} fwd_ms {r.get('decoder_fwd_kernel_ms',0):.3f} step_ms {t['ms_per_step']:.3f} launches {t.get('gpu_launches')} sm_mhz {t['clocks']['sm_mhz']}",
          "fine", round((d.get("train_fine") or {}).get("ms_per_step", 0), 3))
except Exception as e:
    print("ERR", e); print(open(sys.argv[1]).read()[-1500:])
PY
done
done
