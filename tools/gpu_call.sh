mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k01_source_hrma|k1_hrma|k2_select|k2_grating|k3_acis" -s 7 -c 7 -o gpurun_out/prof_r02a python tools/ncu_probe.py 16777216 c2_hetg_acis_s 2 > gpurun_out/c1_ncu.log 2>&1
ls -la gpurun_out/ >> gpurun_out/c1_ncu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-pileup --no-level1 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
tail -c 600 gpurun_out/c1_bench.json
