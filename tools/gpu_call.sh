mkdir -p gpurun_out
for v in default sel default sel; do
if [ $v = default ]; then unset MARXB200_LIB; else export MARXB200_LIB=$PWD/build/variants/libmarxb200_$v.so; fi
timeout 120 python bench.py --steps 60 --warmup 3 --no-sweep --no-probe --no-configs --no-driver --no-cpu-baseline > gpurun_out/bench_v.json 2> /dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_v.json").read().strip().splitlines()[-1])
k=d["roofline"]["kernels"]
print("$v value %.4g ms %.4f | " % (d["value"], d["ms_per_step"]), " ".join("%.4f"%v["ms"] for v in k.values()))
PY
done
unset MARXB200_LIB
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 )
