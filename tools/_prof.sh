set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-framespec-leg --no-disk-leg --no-latency-leg --stream-frames 128"
$CMD > gpurun_out/r02_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu1.log 2>&1
$CMD > gpurun_out/r02_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:k_medoid_screen_sym|k_project_count|k_compact|k_aggregate|k_erode3x3|k_medoid_verify|k_medoid_classify" -s 21 -c 7 -o gpurun_out/r02_prof $CMD > gpurun_out/r02_ncu2.log 2>&1
CMD3="python bench.py --config c3 --steps 2 --warmup 3 --no-cpu-baseline --no-framespec-leg --no-disk-leg --no-latency-leg --stream-frames 64"
$CMD3 > gpurun_out/r02_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:k_hull_obb|k_medoid_screen$" -s 4 -c 2 -o gpurun_out/r02_prof_c3 $CMD3 > gpurun_out/r02_ncu3.log 2>&1
ls -la gpurun_out/
