mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3 )
for rep in 1 2; do
MARXB200_BENCH_HANG_S=240 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$rep bench.py --gpus 2 --steps 60 --warmup 3 --no-sweep --no-probe --no-configs --no-driver --no-cpu-baseline > gpurun_out/bench_n2_rd.json 2> gpurun_out/bench_n2_rd.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n2_rd.json").read().strip().splitlines()[-1])
print("N=2 value %.4g ms %.4f e2e %.4g nomerge %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["merge"]["value_without_merge"]))
PY
done
timeout 120 python bench.py --steps 60 --warmup 3 --no-sweep --no-probe --no-configs --no-driver --no-cpu-baseline > gpurun_out/bench_n1_rd.json 2> /dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n1_rd.json").read().strip().splitlines()[-1])
print("N=1 value %.4g ms %.4f e2e %.4g k01 %.4f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["avg_launch_ms"]))
PY
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 )
