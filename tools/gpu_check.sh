python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "full_size" > gpurun_out/t6.log 2>&1; echo "rc=$?" >> gpurun_out/t6.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
