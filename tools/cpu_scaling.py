"""Scaling of the CPU arm (oracle/batch_oracle.cpp, static OpenMP schedule) over thread counts, and what the box lets a process use.

    python tools/cpu_scaling.py [--frames 256]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--mode", type=int, default=1)
    args = ap.parse_args()
    import bench
    for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us", "/sys/fs/cgroup/cpuset.cpus.effective"):
        if os.path.exists(p):
            print(p, "=", open(p).read().strip())
    print("os.cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "physical", bench.physical_cores(), "cpu", bench.cpu_model())
    try:
        print("loadavg", open("/proc/loadavg").read().strip())
    except OSError:
        pass
    wl = bench.make_workload("c3", 0, frames=min(args.frames, 256))
    n = wl["F"]
    base = None
    for th in (1, 2, 4, 8, 16, 32, 64, 128):
        if th > (os.cpu_count() or 1):
            break
        m = n if th >= 8 else max(4, min(n, 4 * th))
        bench.cpu_run(wl, m, th, line_mode=args.mode)
        dt = min(bench.cpu_run(wl, m, th, line_mode=args.mode)[0] for _ in range(2))
        fps = m / dt
        base = base or fps
        print("threads %3d: %8.1f frames/s  (x%.1f of one thread, %d frames)" % (th, fps, fps / base, m), flush=True)


if __name__ == "__main__":
    main()
