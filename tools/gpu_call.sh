mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_zz_pileup.py tests/test_gpu_marx_driver.py -x -q -m gpu 2>&1 | tail -25 ) > gpurun_out/c5_tests_a.log
cat gpurun_out/c5_tests_a.log
( timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_zz_pileup.py --deselect tests/test_gpu_marx_driver.py 2>&1 | tail -15 ) > gpurun_out/c5_tests_b.log
cat gpurun_out/c5_tests_b.log
( MARXB200_BENCH_TRACE=1 timeout 600 python bench.py --steps 20 --warmup 3 ) > gpurun_out/c5_bench_n1.json 2> gpurun_out/c5_bench_n1.err
tail -c 2500 gpurun_out/c5_bench_n1.json; tail -12 gpurun_out/c5_bench_n1.err
