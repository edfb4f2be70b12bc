set -x
cd /root/repo
mkdir -p gpurun_out
timeout 400 python bench.py 2>&1 | tail -1 > gpurun_out/r01_final5_bench_line.json
cut -c1-300 gpurun_out/r01_final5_bench_line.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/r01_final5_reference_line.json
cut -c1-400 gpurun_out/r01_final5_reference_line.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 130 --csv --log-file gpurun_out/r01_final5_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k0_time|k1_hrma|k01_source_hrma|k3_acis|k2_grating|order_|l1_" -s 19 -c 19 -f -o gpurun_out/prof_r01_final5 python tools/ncu_probe.py 16777216 c2_hetg_acis_s 2 2>&1 | tail -2
