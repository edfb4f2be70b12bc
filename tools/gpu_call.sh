cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/call55_k2sel_variants.log
: > $L
for v in default MB8 ILP2 ILP2MB6; do
  lib=/root/repo/build/variants/libmarxb200_$v.so
  [ "$v" = default ] && lib=/root/repo/marx_b200/libmarxb200.so
  for cfg in c2_hetg_acis_s c3_letg_hrc_s; do
    MARXB200_LIB=$lib timeout 60 python tools/trace_probe.py 16777216 $cfg 6 2>&1 | cut -c1-330 >> $L
  done
done
for v in ILP2 ILP2MB6; do
  ( MARXB200_LIB=/root/repo/build/variants/libmarxb200_$v.so timeout 120 python -m pytest tests/test_gpu_golden.py tests/test_gpu_oracle.py -q -m gpu -x -k "compacted" 2>&1 | tail -3 ) >> gpurun_out/call55_tests.log
done
cat gpurun_out/call55_tests.log
