# interleaved A/B of library variants on the training step, printing the final loss too: VARIANTS="a b" ROUNDS=2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,serial --format=csv,noheader
for r in $(seq 1 ${ROUNDS:-2}); do
for v in default ${VARIANTS}; do
  if [ $v = default ]; then unset UORF_B200_LIB; else export UORF_B200_LIB=$PWD/uorf_b200/_build/variants/libuorf_b200_$v.so; fi
  timeout 900 python bench.py --mode train --steps ${STEPS:-15} --warmup 4 --no-cpu-baseline > gpurun_out/ab7_${v}.log 2>&1
  python - gpurun_out/ab7_${v}.log $v <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    t=d.get("train") or d
    r=t.get("roofline") or {}
    print(f"{sys.argv[2]:14s} bwd_ms {r.get('kernel_ms',0):.3f} fwd_ms {r.get('decoder_fwd_kernel_ms',0):.3f} step_ms {t['ms_per_step']:.3f} loss {t.get('loss_last')} sm_mhz {t['clocks']['sm_mhz']}")
except Exception as e:
    print("ERR", e); print(open(sys.argv[1]).read()[-800:])
PY
done
done
