mkdir -p gpurun_out
timeout 120 python tools/pileup_ncu_probe.py 2>&1 | tail -1
( timeout 300 python -m pytest tests/test_gpu_zz_pileup.py -x -q -m gpu 2>&1 | tail -3 )
