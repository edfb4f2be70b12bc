set -x
cd /root/repo
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6
python bench.py --steps 100 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r01_final2.json; cut -c1-300 gpurun_out/bench_r01_final2.json
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_r01_reference.json; cat gpurun_out/bench_r01_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r01_final2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k1_hrma|k01_source_hrma|k3_acis|k2_grating|order_|k0_time" -s 10 -c 12 -o gpurun_out/prof_r01_final2 python tools/ncu_probe.py 16777216 c2_hetg_acis_s 2 2>&1 | tail -3
