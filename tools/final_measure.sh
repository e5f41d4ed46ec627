#!/bin/bash
# Round-end measurement pass on one B200 (run under gpurun): GPU tests, the bench lines of every config, the
# reference arm, then - each only after the same command exited 0 WITHOUT ncu - the launch lists (C2, C3) and one
# full-metric capture of the top kernels.  Usage: bash tools/final_measure.sh <prefix>   (files: gpurun_out/<prefix>_*)
P=${1:-r02c}
O=gpurun_out
set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/${P}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/${P}_smoke.log 2>&1
timeout 700 python bench.py --steps 20 --warmup 5 > $O/${P}_c2.json 2> $O/${P}_c2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${P}_ref.json 2> $O/${P}_ref.err
for c in c1 c3 c4; do timeout 600 python bench.py --config $c --steps 20 --warmup 5 > $O/${P}_$c.json 2> $O/${P}_$c.err; done
LEGS="--no-cpu-baseline --no-framespec-leg --no-disk-leg --no-latency-leg"
CMD="python bench.py --steps 2 --warmup 3 $LEGS --stream-frames 128 --e2e-ramp 0"
$CMD > $O/${P}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file $O/${P}_launches.csv $CMD > $O/${P}_ncu1.log 2>&1
CMD3="python bench.py --config c3 --steps 2 --warmup 3 $LEGS --stream-frames 64 --e2e-ramp 0"
$CMD3 > $O/${P}_plain3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 500 --csv --log-file $O/${P}_c3_launches.csv $CMD3 > $O/${P}_ncu3.log 2>&1
if [ "$2" = "full" ]; then
$CMD > $O/${P}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:k_medoid_screen_sym|k_project_count|k_compact|k_aggregate|k_erode3x3" -s 15 -c 5 -o $O/${P}_prof $CMD > $O/${P}_ncu2.log 2>&1
fi
ls -la $O | tail -20
