cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_level1.py -q -m gpu -x 2>&1 | tail -25 ) | tee gpurun_out/call47_tests.log
