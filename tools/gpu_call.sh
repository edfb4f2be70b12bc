set -x
cd /root/repo
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_level1.py -q -m gpu -x 2>&1 | tail -25 ) | tee gpurun_out/call29_tests.log
( timeout 120 python tools/level1_probe.py 16777216 c2_hetg_acis_s level1_acis_s_hetg_edser;
  timeout 120 python tools/level1_probe.py 16777216 c1_acis_s level1_acis_s_nodither_none;
  timeout 120 python tools/level1_probe.py 16777216 c1_acis_s level1_acis_s_hetg_edser ) 2>&1 | grep -v "^+" | tee gpurun_out/call29_probe.log
