cd /root/repo
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_marx_driver.py -q -m gpu -x -k "user_source or rayfile" 2>&1 | tail -25 ) | tee gpurun_out/call51_tests.log
