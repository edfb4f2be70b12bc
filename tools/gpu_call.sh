set -x
cd /root/repo
timeout 1700 python -m pytest tests -q -m gpu 2>&1 | tail -6
python bench.py --steps 100 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r01_final3.json; cut -c1-200 gpurun_out/bench_r01_final3.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 90 --csv --log-file gpurun_out/launches_r01_final3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k1_hrma|k01_source_hrma|k3_acis|k2_grating|order_|k0_time" -s 15 -c 15 -o gpurun_out/prof_r01_final3 python tools/ncu_probe.py 16777216 c2_hetg_acis_s 2 2>&1 | tail -3
