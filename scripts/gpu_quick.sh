# quick GPU visit: parity tests + bench in the given precisions (no profiler)
mkdir -p gpurun_out; rm -f gpurun_out/*.log
for grp in ${TEST_GROUPS:-probe composite sample_coords slot_attention decoder_golden decoder_vs_oracle eval_driver}; do
  echo "=== $grp" >> gpurun_out/tests.log
  timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s --timeout 240 -k "$grp" 2>&1 | tail -${TAILN:-45} >> gpurun_out/tests.log
done
python -c "
from uorf_b200 import _lib
print('debug flag', _lib.lib().uorf_debug_flag())" >> gpurun_out/tests.log 2>&1
for prec in ${PRECS:-fp16x3 bf16}; do
  timeout 600 python bench.py --steps 30 --warmup 5 --precision $prec --no-cpu-baseline > gpurun_out/bench_$prec.log 2>&1
done
grep -E "passed|failed|===|Error|debug flag" gpurun_out/tests.log; for f in gpurun_out/bench_*.log; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "dec_ms", d["roofline"]["kernel_ms"], "frac", round(d["roofline"]["frac"],4), d["clocks"])
except Exception as e:
    print("ERR", e); print(open(sys.argv[1]).read()[-1500:])
PY
done
if [ -n "$TRAIN" ]; then
  timeout 900 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/bench_train.log 2>&1; tail -c 1800 gpurun_out/bench_train.log
fi
