cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 ) | tee gpurun_out/call54_tests.log
timeout 600 python bench.py > gpurun_out/r01_final7_bench_line.json 2> gpurun_out/call54_bench.err
tail -c 300 gpurun_out/r01_final7_bench_line.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r01_final7_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/call54_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k0_time|k1_hrma|k01_source_hrma|k3_acis|k2_|order_|l1_" \
    -s 21 -c 21 -f -o gpurun_out/prof_r01_final7 python tools/ncu_probe.py 16777216 c2_hetg_acis_s 2 > gpurun_out/call54_ncu_full.log 2>&1
tail -3 gpurun_out/call54_ncu_full.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) | tee gpurun_out/call54_smoke.log
for cfg in c1_acis_s c4_beta_acis_i; do timeout 300 python tools/trace_probe.py 16777216 $cfg 10 >> gpurun_out/call54_other.log 2>&1; done
