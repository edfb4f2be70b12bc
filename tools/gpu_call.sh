set -x
cd /root/repo
timeout 1500 python -m pytest tests/test_gpu_edges.py tests/test_gpu_golden.py -q -m gpu 2>&1 | tail -25
timeout 300 python tools/trace_probe.py 16777216 c2_hetg_acis_s 10 2>&1 | tail -1
