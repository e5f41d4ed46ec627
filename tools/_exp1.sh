set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02k_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-disk-leg --no-latency-leg > gpurun_out/r02k_a.json 2> gpurun_out/r02k_a.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-disk-leg --no-latency-leg --no-overlap --no-framespec-leg > gpurun_out/r02k_b.json 2> gpurun_out/r02k_b.err
