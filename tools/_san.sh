cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/sanitize_driver.py > gpurun_out/san_plain.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/san_plain.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 --print-limit 20 python tools/sanitize_driver.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/r02_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 --print-limit 20 python tools/sanitize_driver.py --small > gpurun_out/r02_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/r02_sanitizer_racecheck.log
