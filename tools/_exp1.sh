set -x
for pad in 0 29 40; do
CM3D_SYM_PAD_KB=$pad timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-disk-leg --no-latency-leg --no-framespec-leg --stream-frames 512 > gpurun_out/r03b_$pad.json 2> gpurun_out/r03b_$pad.err
done
