set -x
cd /root/repo
timeout 1700 python -m pytest tests -q -m gpu -x 2>&1 | tail -12
for c in c2_hetg_acis_s c1_acis_s c3_letg_hrc_s; do timeout 300 python tools/trace_probe.py 16777216 $c 20 2>&1 | tail -1; done
