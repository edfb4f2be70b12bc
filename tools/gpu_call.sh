mkdir -p gpurun_out
( nvidia-smi topo -m; echo; ls /sys/devices/system/node/ 2>&1 | head; echo; for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302\|^0x0300" $d/class 2>/dev/null; then echo "$d numa=$(cat $d/numa_node 2>/dev/null) local_cpus=$(cat $d/local_cpulist 2>/dev/null)"; fi; done; echo; grep -i "allowed_list" /proc/self/status; nproc; lscpu | grep -i "numa\|socket\|model name" ; python -c "
import ctypes, os
libc=ctypes.CDLL('libc.so.6', use_errno=True)
import ctypes as C
# get_mempolicy syscall 239 on x86_64
mode=C.c_int(); mask=(C.c_ulong*16)()
r=libc.syscall(239, C.byref(mode), mask, 1024, None, 0)
print('get_mempolicy rc', r, 'errno', C.get_errno(), 'mode', mode.value)
" ) > gpurun_out/topology.txt 2>&1
cat gpurun_out/topology.txt | cut -c1-220 | head -60
timeout 120 python bench.py --steps 60 --warmup 3 --no-sweep --no-probe --no-configs --no-driver --no-cpu-baseline > gpurun_out/bench_n1_x.json 2> /dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n1_x.json").read().strip().splitlines()[-1])
print("N=1 value %.4g ms %.4f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
for k,v in d["roofline"]["kernels"].items(): print("   %-45s %.4f ms" % (k, v["ms"]))
PY
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 )
