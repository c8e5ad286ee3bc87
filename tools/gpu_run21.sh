set -x
mkdir -p gpurun_out
timeout 150 compute-sanitizer --tool racecheck --print-limit 30 python tools/sanitize_probe.py 96 > gpurun_out/racecheck.log 2>&1; grep -c "Race reported\|hazard" gpurun_out/racecheck.log; grep -A6 "hazard\|Race reported" gpurun_out/racecheck.log | head -60; tail -4 gpurun_out/racecheck.log
timeout 150 compute-sanitizer --tool initcheck --print-limit 30 python tools/sanitize_probe.py 96 > gpurun_out/initcheck.log 2>&1; grep -A8 "Uninitialized" gpurun_out/initcheck.log | head -60; tail -4 gpurun_out/initcheck.log
