#!/bin/bash
# One GPU-box session: parity tests, LSD seed-loop timing, benches at several numbers of batches in flight.  Outputs under gpurun_out/<tag>_*.
TAG=${1:-t}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${TAG}_tests.log
timeout 200 python tools/time_lines.py --flavour lsd --reps 3 > gpurun_out/${TAG}_lines.log 2>&1
cat gpurun_out/${TAG}_lines.log
for nf in ${2:-12 16 20}; do
  timeout 600 python bench.py --steps 48 --warmup 3 --inflight $nf --no-configs > gpurun_out/${TAG}_bench_if$nf.json 2> gpurun_out/${TAG}_bench_if$nf.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_if$nf.json").read().strip().splitlines()[-1])
print("inflight $nf: ms/step", d["ms_per_step"], "frames/s", d["frames_per_s"], "e2e ms", d["e2e"]["ms_per_step"], "e2e frames/s", d["e2e"]["frames_per_s"], "edl", d.get("online_edlines",{}).get("ms_per_step"))
PY
  tail -3 gpurun_out/${TAG}_bench_if$nf.err
done
