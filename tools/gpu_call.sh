set -x
cd /root/repo
timeout 1200 python -m pytest tests/test_gpu_marx_driver.py -q -m gpu -x 2>&1 | tail -25
