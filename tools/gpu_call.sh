# scratch command file of the development loop: `gpurun -- 'bash tools/gpu_call.sh'` (see profiles/README.md for the commands behind the
# committed evidence).  Default: the GPU suite and one short bench line.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 )
timeout 200 python bench.py --steps 60 --warmup 3 --no-sweep --no-probe --no-configs --no-driver --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
cut -c1-300 gpurun_out/bench_quick.json
