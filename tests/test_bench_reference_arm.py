"""`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) end to end on this machine: it must print exactly one JSON
line on stdout with the contract's keys, time the same path (line detection + detect_cuboid per frame), and say which code ran stage (i)
-- the reference's own compiled detect_filter_lines where oracle/_ref exists, else the oracle port."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line(oracle):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:500]          # the reference's own chatter on std::cout is switched off in oracle/ref/linelbd_ref.cpp
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "scored cuboid proposals/s" and d["unit"] == "proposals/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["value"] == d["value"] and cb["cores"] >= 1 and "frames" in cb["sample"]
    want = "reference" if oracle.ref_detect_filter_lines_available() else "port"
    assert cb["kind"] == want
    assert 150 < d["config"]["segments_per_frame_M"] < 400        # the LSD flavour really ran on the frames (section 8(d) density)
