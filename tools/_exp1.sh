set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02t_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-disk-leg --no-latency-leg --config c4 > gpurun_out/r02t_c4.json 2> gpurun_out/r02t_c4.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-disk-leg --no-latency-leg > gpurun_out/r02t_c2.json 2> gpurun_out/r02t_c2.err
