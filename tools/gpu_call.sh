mkdir -p gpurun_out
export NCCL_DEBUG=WARN MARXB200_BENCH_TRACE=1 MARXB200_BENCH_HANG_S=120
( timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 10 --warmup 3 --sweep-max 1e8 ) > gpurun_out/c3_bench_n2.json 2> gpurun_out/c3_bench_n2.err
echo "rc=$?"; tail -c 1200 gpurun_out/c3_bench_n2.json; grep -E "bench rank|Error|error|Traceback|File \"/root|line [0-9]+ in" gpurun_out/c3_bench_n2.err | tail -60
