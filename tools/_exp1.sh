set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02w_tests.log
for a in "32 0 c3" "32 2000 c3" "16 0 c4" "16 2000 c4"; do timeout 600 python tools/validate_screen.py $a >> gpurun_out/r02w_validate.txt 2>&1; done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-disk-leg --no-latency-leg --no-framespec-leg --config c3 > gpurun_out/r02w_c3.json 2> gpurun_out/r02w_c3.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-disk-leg --no-latency-leg --no-framespec-leg --config c4 > gpurun_out/r02w_c4.json 2> gpurun_out/r02w_c4.err
