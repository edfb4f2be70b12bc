set -x
cd /root/repo
timeout 1700 python -m pytest tests -q -m gpu 2>&1 | tail -6
timeout 300 python tools/trace_probe.py 16777216 c2_hetg_acis_s 20 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k01_source_hrma" -s 1 -c 1 python tools/ncu_probe.py 16777216 c2_hetg_acis_s 2 2>&1 | grep -E "dram__|gpu__time" 
