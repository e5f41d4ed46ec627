set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_pass2_functions.py tests/test_ref_script_goldens.py -m gpu -x -q -k "hull or kitti or obb" 2>&1 | tail -15 > gpurun_out/r02o_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-disk-leg --no-framespec-leg --no-latency-leg --config c3 > gpurun_out/r02o_c3.json 2> gpurun_out/r02o_c3.err
