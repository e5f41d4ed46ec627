set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "graph or c5_seeds" 2>&1 | tail -15 > gpurun_out/r02m_tests.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-disk-leg --no-framespec-leg --config c1 > gpurun_out/r02m_c1.json 2> gpurun_out/r02m_c1.err
