set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r03a_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-disk-leg --no-latency-leg --config c4 > gpurun_out/r03a_c4.json 2> gpurun_out/r03a_c4.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-disk-leg --no-latency-leg --config c1 > gpurun_out/r03a_c1.json 2> gpurun_out/r03a_c1.err
