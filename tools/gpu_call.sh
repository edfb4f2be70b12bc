# scratch command file of the development loop: `gpurun -- 'bash tools/gpu_call.sh'` (see profiles/README.md for the commands behind the
# committed evidence).  Default: the GPU suite and one short bench line.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 )
timeout 200 python bench.py --steps 60 --warmup 3 --no-sweep --no-probe --no-configs --no-driver --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
k=d["roofline"]["kernels"]
print("value %.4g ms %.4f e2e %.4g | " % (d["value"], d["ms_per_step"], d["e2e"]["value"]), " ".join("%.4f"%v["ms"] for v in k.values()))
PY
