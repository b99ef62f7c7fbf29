# compute-sanitizer gate, ONE tool per gpurun call (B200_PROFILING.md):  bash scripts/gpu_sanitize.sh memcheck|racecheck|synccheck|initcheck
tool=${1:-memcheck}
mkdir -p gpurun_out
timeout 300 python scripts/sanitize_case.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/sanitize_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 python scripts/sanitize_case.py > gpurun_out/sanitize_$tool.log 2>&1
echo "compute-sanitizer --tool $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_case: ok|Error|error" gpurun_out/sanitize_$tool.log | head -20
