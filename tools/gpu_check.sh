#!/bin/bash
# 2-GPU check: the drop-in scripts under torchrun give byte-identical labels to a single process
set -x
rm -rf /tmp/cm3d_a /tmp/cm3d_b
python tools/run_synthetic_scripts.py --out /tmp/cm3d_a > gpurun_out/scripts_1p.log 2>&1; echo "rc1=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/run_synthetic_scripts.py --out /tmp/cm3d_b > gpurun_out/scripts_2p.log 2>&1; echo "rc2=$?"
diff -r -x '_data_rank*' /tmp/cm3d_a /tmp/cm3d_b > gpurun_out/scripts_diff.log 2>&1; echo "diff_rc=$?"
find /tmp/cm3d_a -type f -not -path '*_data_rank*' | sort | xargs ls -la >> gpurun_out/scripts_diff.log
