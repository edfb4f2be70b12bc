set -x
cd /root/repo
timeout 1500 python -m pytest tests/test_gpu_param_surface.py -q -m gpu 2>&1 | tail -25
