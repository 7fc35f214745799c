#!/bin/bash
# First GPU session of the code written after round 2's GPU minutes were spent (last bullet): the descriptor / matcher
# kernels (cs_lbd.cu) and the carried-pose mode (cs_set_profiling bit 10).  One gpurun call:
#     gpurun --timeout 900 -- 'bash tools/gpu_lbd_first_run.sh r3a'
# Outputs under gpurun_out/<tag>_*: the parity tests of that code (the file's own order: everything that does not depend on the EDLines kernels'
# two new stores first), the timings, an ncu launch list of the timing run and one full capture of k_lbd_describe.
TAG=${1:-lbd}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_z_gpu_lbd_parity.py -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1; echo "z tests rc=$?"; tail -6 gpurun_out/${TAG}_tests.log
timeout 300 python tools/time_lbd.py --frames 256 > gpurun_out/${TAG}_time_lbd.json 2> gpurun_out/${TAG}_time_lbd.err; echo "time_lbd rc=$?"; cat gpurun_out/${TAG}_time_lbd.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_lbd.csv python tools/time_lbd.py --frames 64 --reps 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_lbd_describe -c 1 -o gpurun_out/${TAG}_lbd_describe python tools/time_lbd.py --frames 64 --reps 1 > /dev/null 2>&1
ls -la gpurun_out | tail -8
