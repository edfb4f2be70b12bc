set -x
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 | tee gpurun_out/call25.log
timeout 600 python bench.py > gpurun_out/bench_tmp.json 2> gpurun_out/b.log; tail -c 600 gpurun_out/bench_tmp.json
