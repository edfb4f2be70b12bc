mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_zz_pileup.py -x -q -m gpu 2>&1 | tail -12 ) > gpurun_out/c8_tests_pu.log; cat gpurun_out/c8_tests_pu.log
timeout 120 python tools/pileup_ncu_probe.py 2>&1 | tail -2
( timeout 900 python -m pytest tests/test_gpu_statistics.py tests/test_gpu_param_surface.py tests/test_gpu_golden.py tests/test_gpu_oracle.py -x -q -s -m gpu -k "distributions or flatfield or ray_index or golden or slot_by_slot" 2>&1 | grep -v "^$" | tail -40 | cut -c1-1200 ) > gpurun_out/c8_tests.log
cat gpurun_out/c8_tests.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"pu_fused" -s 1 -c 1 -o gpurun_out/prof_r02_pileup2 python tools/pileup_ncu_probe.py > gpurun_out/c8_ncu_pu.log 2>&1; tail -2 gpurun_out/c8_ncu_pu.log
