mkdir -p gpurun_out
timeout 120 python bench.py --steps 60 --warmup 3 --no-sweep --no-probe --no-configs --no-driver --no-cpu-baseline > gpurun_out/bench_n1_x.json 2> /dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n1_x.json").read().strip().splitlines()[-1])
print("N=1 value %.4g ms %.4f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
for k,v in d["roofline"]["kernels"].items(): print("   %-45s %.4f ms" % (k, v["ms"]))
PY
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 )
