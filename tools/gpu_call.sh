cd /root/repo
mkdir -p gpurun_out
for v in base cur base cur; do
  if [ $v = cur ]; then L=/root/repo/marx_b200/libmarxb200.so; else L=/root/repo/build/variants/libmarxb200_$v.so; fi
  MARXB200_LIB=$L timeout 120 python tools/trace_probe.py 16777216 c2_hetg_acis_s 30 2>&1 | tail -1
done | tee gpurun_out/call40_variants.log
( timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_oracle.py -q -m gpu -x 2>&1 | tail -4 ) | tee gpurun_out/call40_tests.log
