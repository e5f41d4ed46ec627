set -x
for pw in 6 8 12; do
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-disk-leg --no-latency-leg --pack-workers $pw > gpurun_out/r02j_$pw.json 2> gpurun_out/r02j_$pw.err
done
