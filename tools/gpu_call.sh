cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/call53_k2split.log
: > $L
for cfg in c2_hetg_acis_s c3_letg_hrc_s; do
  for sp in 1 0; do
    echo "== $cfg K2_SPLIT=$sp" >> $L
    MARXB200_K2_SPLIT=$sp timeout 300 python tools/trace_probe.py 16777216 $cfg 10 >> $L 2>&1
  done
done
( timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_oracle.py tests/test_gpu_edges.py tests/test_gpu_param_surface.py -q -m gpu -x 2>&1 | tail -25 ) | tee gpurun_out/call53_tests.log
