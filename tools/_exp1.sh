set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02p_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-disk-leg --no-framespec-leg --no-latency-leg > gpurun_out/r02p_c2.json 2> gpurun_out/r02p_c2.err
