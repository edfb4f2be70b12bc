mkdir -p gpurun_out
for w in 256 512 1024; do echo "window $w: $(MARXB200_PILEUP_WINDOW=$w timeout 120 python tools/pileup_ncu_probe.py 2>&1 | tail -1)"; done
for w in 256 512; do ( MARXB200_PILEUP_WINDOW=$w timeout 300 python -m pytest tests/test_gpu_zz_pileup.py -x -q -m gpu 2>&1 | tail -2 ); done
