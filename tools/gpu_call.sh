mkdir -p gpurun_out
MARXB200_BENCH_HANG_S=300 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 3 --no-configs --no-driver --no-cpu-baseline > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n2.json").read().strip().splitlines()[-1])
print("N=2 value %.4g ms %.4f e2e %.4g nomerge %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["merge"]["value_without_merge"]))
PY
timeout 120 python bench.py --steps 60 --warmup 3 --no-sweep --no-probe --no-configs --no-driver --no-cpu-baseline > gpurun_out/r02_bench_n1_samebox_as_n2.json 2> /dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n1_samebox_as_n2.json").read().strip().splitlines()[-1])
print("N=1 same box value %.4g ms %.4f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
PY
