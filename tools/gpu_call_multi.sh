set -x
cd /root/repo
N=${1:-2}
nvidia-smi -L | head -8
if [ "$N" = "2" ]; then timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -4; fi
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 30 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_r01_final7_n$N.json
cut -c1-400 gpurun_out/bench_r01_final7_n$N.json
