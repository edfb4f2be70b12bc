cd /root/repo
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_marxasp.py -q -m gpu -x -s 2>&1 | tail -30 ) | tee gpurun_out/call52_tests.log
