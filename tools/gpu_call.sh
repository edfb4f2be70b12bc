( timeout 600 python -m pytest tests/test_gpu_lookahead.py -x -q -m gpu 2>&1 | tail -12 | cut -c1-300 )
