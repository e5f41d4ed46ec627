set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r02u_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-disk-leg --no-latency-leg --no-framespec-leg --config c3 > gpurun_out/r02u_c3.json 2> gpurun_out/r02u_c3.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-disk-leg --no-latency-leg --no-framespec-leg --config c4 > gpurun_out/r02u_c4.json 2> gpurun_out/r02u_c4.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-disk-leg --no-latency-leg --no-framespec-leg > gpurun_out/r02u_c2.json 2> gpurun_out/r02u_c2.err
