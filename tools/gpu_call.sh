cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 ) | tee gpurun_out/call43_tests.log
