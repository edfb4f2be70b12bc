set -x
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_aspsol.py tests/test_gpu_tally.py tests/test_gpu_marx_driver.py tests/test_gpu_edges.py -q -m gpu 2>&1 | tail -15 | tee gpurun_out/call26.log
for v in base r64all r64k1 sparse2 k01r2 base; do
  MARXB200_LIB=/root/repo/build/variants/libmarxb200_$v.so timeout 120 python tools/trace_probe.py 16777216 c2_hetg_acis_s 20 2>&1 | tail -1
done | tee gpurun_out/variants26.log
