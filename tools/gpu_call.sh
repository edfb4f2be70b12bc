mkdir -p gpurun_out
MARXB200_BENCH_HANG_S=300 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 60 --warmup 3 --no-configs --no-driver --no-cpu-baseline > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
tail -2 gpurun_out/r02_bench_n8.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n8.json").read().strip().splitlines()[-1])
print("N=8 value %.4g ms %.4f e2e %.4g nomerge %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["merge"]["value_without_merge"]))
print(json.dumps(d["e2e"].get("d2h_floor")))
PY
timeout 120 python bench.py --steps 60 --warmup 3 --no-sweep --no-probe --no-configs --no-driver --no-cpu-baseline > gpurun_out/r02_bench_n1_samebox_as_n8.json 2> /dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n1_samebox_as_n8.json").read().strip().splitlines()[-1])
print("N=1 same box value %.4g ms %.4f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
PY
