mkdir -p gpurun_out
for rep in 1 2 3; do for la in 1 0; do
MARXB200_VERBOSE=1 MARXB200_LOOKAHEAD=$la timeout 200 python bench.py --steps 60 --warmup 3 --no-sweep --no-probe --no-configs --no-driver --no-cpu-baseline > gpurun_out/bench_n1_la$la.json 2> gpurun_out/bench_n1_la$la.err
grep "look-ahead" gpurun_out/bench_n1_la$la.err | head -2
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n1_la$la.json").read().strip().splitlines()[-1])
print("N=1 lookahead=$la value %.4g ms %.4f e2e %.4g profiled ms %.4f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["profiled_ms_per_step"]))
PY
done; done
