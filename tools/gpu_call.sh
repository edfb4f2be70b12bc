set -x
cd /root/repo
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests/test_gpu_marx_driver.py -x -q -m gpu 2>&1 | tail -15
for v in "" outl fmad fmad_outl; do
  if [ -z "$v" ]; then unset MARXB200_LIB; else export MARXB200_LIB=/root/repo/build/variants/libmarxb200_$v.so; fi
  echo "=== variant [$v]"
  timeout 300 python tools/perf_probe.py 16777216 c2_hetg_acis_s 4 2>&1 | tail -2
  timeout 300 python tools/trace_probe.py 16777216 c2_hetg_acis_s 10 2>&1 | tail -1
done
export MARXB200_LIB=/root/repo/build/variants/libmarxb200_fmad.so
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_oracle.py -q -m gpu 2>&1 | tail -15
unset MARXB200_LIB
timeout 1200 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_marx_driver.py 2>&1 | tail -5
