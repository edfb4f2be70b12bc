mkdir -p gpurun_out
MARXB200_BENCH_HANG_S=300 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 4 --steps 60 --warmup 3 --no-configs --no-driver --no-cpu-baseline > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err
tail -1 gpurun_out/r02_bench_n4.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n4.json").read().strip().splitlines()[-1])
print("N=4 value %.4g ms %.4f e2e %.4g nomerge %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["merge"]["value_without_merge"]))
print(json.dumps(d["e2e"].get("d2h_floor"))[:300]); print(d["e2e"]["d2h_ceiling"]["concurrent_gbs_per_rank"])
PY
timeout 120 python bench.py --steps 60 --warmup 3 --no-sweep --no-probe --no-configs --no-driver --no-cpu-baseline > gpurun_out/r02_bench_n1_samebox_as_n4.json 2> /dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n1_samebox_as_n4.json").read().strip().splitlines()[-1])
print("N=1 same box value %.4g ms %.4f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
PY
