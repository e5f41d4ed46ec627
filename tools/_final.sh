set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02r_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02r_c2.json 2> gpurun_out/r02r_c2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02r_ref.json 2> gpurun_out/r02r_ref.err
for c in c1 c3 c4; do timeout 600 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r02r_$c.json 2> gpurun_out/r02r_$c.err; done
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-framespec-leg --no-disk-leg --no-latency-leg --stream-frames 128 --e2e-ramp 0"
$CMD > gpurun_out/r02r_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02r_launches.csv $CMD > gpurun_out/r02r_ncu1.log 2>&1
