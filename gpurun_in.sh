timeout 120 python tools/sanitize_probe.py 2>&1 | tail -16
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_probe.py > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|Invalid|MISMATCH" gpurun_out/memcheck.log | head -10
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_probe.py > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|hazard|MISMATCH" gpurun_out/racecheck.log | head -10
