#!/bin/bash
# One GPU-box session: parity tests, line-detector timings, a short bench.  Outputs under gpurun_out/<tag>_*.
TAG=${1:-t}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${TAG}_tests.log
timeout 200 python tools/time_lines.py --flavour lsd > gpurun_out/${TAG}_lines.log 2>&1
timeout 200 python tools/time_lines.py --flavour lsd --seq --reps 2 >> gpurun_out/${TAG}_lines.log 2>&1
timeout 200 python tools/time_lines.py --flavour edlines --reps 3 >> gpurun_out/${TAG}_lines.log 2>&1
timeout 200 python tools/time_lines.py --flavour edlines --seq --reps 2 >> gpurun_out/${TAG}_lines.log 2>&1
cat gpurun_out/${TAG}_lines.log
timeout 600 python bench.py --steps 50 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 7000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
