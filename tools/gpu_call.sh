set -x
cd /root/repo
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k3_acis|k01_source_hrma" -s 2 -c 2 -o gpurun_out/prof_r01_k3 python tools/ncu_probe.py 16777216 c2_hetg_acis_s 2 2>&1 | tail -3
