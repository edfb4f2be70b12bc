set -x
cd /root/repo
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_probe.py 20000 > gpurun_out/san_mem.log 2>&1; echo "memcheck rc=$?"; tail -12 gpurun_out/san_mem.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_probe.py 6000 > gpurun_out/san_race.log 2>&1; echo "racecheck rc=$?"; tail -8 gpurun_out/san_race.log
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 9 python tools/sanitize_probe.py 6000 > gpurun_out/san_init.log 2>&1; echo "initcheck rc=$?"; tail -8 gpurun_out/san_init.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize_probe.py 6000 > gpurun_out/san_sync.log 2>&1; echo "synccheck rc=$?"; tail -5 gpurun_out/san_sync.log
