set -x
cd /root/repo
for v in "" pf pf2 pf4 c4; do
  if [ -z "$v" ]; then unset MARXB200_LIB; else export MARXB200_LIB=/root/repo/build/variants/libmarxb200_$v.so; fi
  echo "=== variant [$v]"
  timeout 300 python tools/trace_probe.py 16777216 c2_hetg_acis_s 20 2>&1 | tail -1
done
