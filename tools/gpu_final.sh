#!/bin/bash
# Round-end evidence in one GPU-box session: parity tests, smoke, the bench line (both arms), ncu launch lists.  Outputs: gpurun_out/<tag>_*
TAG=${1:-fin}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${TAG}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 48 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
r = json.loads(open("gpurun_out/${TAG}_bench_ref.json").read().strip().splitlines()[-1])
print("ours: ms/step %.3f value %.0f frames/s %.0f | e2e ms %.3f value %.0f | launches %s | clocks %s" % (d["ms_per_step"], d["value"], d["frames_per_s"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["clocks"]))
print("cpu_baseline:", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"].get("lsd_stage_one_core"))
print("reference arm: value %.0f frames/s %.0f cores %s" % (r["value"], r["frames_per_s"], r["cpu_baseline"]["cores"]))
print("roofline:", {k: d["roofline"][k] for k in ("kernel", "achieved", "peak", "frac")})
for k in ("lines_given", "online_edlines", "c4", "c5"):
    if k in d: print(k, d[k]["ms_per_step"], d[k]["value"])
PY
for fl in lsd edlines; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_$fl.csv python tools/time_lines.py --flavour $fl --reps 1 > /dev/null 2>&1
done
ls -la gpurun_out | tail -12
