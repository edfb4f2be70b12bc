mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize_probe.py 20000 > gpurun_out/r02_racecheck.txt 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/r02_racecheck.txt | cut -c1-300; grep -c "Race reported\|hazard" gpurun_out/r02_racecheck.txt
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 1 python tools/sanitize_probe.py 20000 > gpurun_out/r02_synccheck.txt 2>&1
echo "synccheck rc=$?"; tail -2 gpurun_out/r02_synccheck.txt | cut -c1-300
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 1 python tools/pileup_ncu_probe.py 200000 > gpurun_out/r02_racecheck_pileup.txt 2>&1
echo "racecheck pileup rc=$?"; tail -3 gpurun_out/r02_racecheck_pileup.txt | cut -c1-300
