#!/usr/bin/env python3
"""Run a command and sample its memory use (RSS, per-GPU used MiB) into a CSV.

Parity: reference ``src/mem_monitor.py`` (:21-159): spawn the command, poll every ``-i`` seconds —
process-tree RSS through psutil, GPU memory (GPUtil there; NVML via ``pynvml``/``nvidia-smi``
here), optional plot.  Jetson ``jtop`` sampling is replaced by NVML (the target is a B200 box).
"""
from __future__ import annotations

import argparse
import csv
import shlex
import subprocess
import time
from pathlib import Path
from typing import List


def gpu_used_mib() -> List[float]:
    try:
        import pynvml

        pynvml.nvmlInit()
        out = []
        for i in range(pynvml.nvmlDeviceGetCount()):
            h = pynvml.nvmlDeviceGetHandleByIndex(i)
            out.append(pynvml.nvmlDeviceGetMemoryInfo(h).used / 2 ** 20)
        return out
    except Exception:  # noqa: BLE001
        try:
            r = subprocess.run(["nvidia-smi", "--query-gpu=memory.used", "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=5)
            return [float(x) for x in r.stdout.split()] if r.returncode == 0 else []
        except Exception:  # noqa: BLE001
            return []


def tree_rss_mib(proc) -> float:
    import psutil

    try:
        p = psutil.Process(proc.pid)
        procs = [p] + p.children(recursive=True)
        return sum(q.memory_info().rss for q in procs if q.is_running()) / 2 ** 20
    except psutil.Error:
        return 0.0


def main(argv=None) -> int:
    p = argparse.ArgumentParser(description=__doc__)
    p.add_argument("command", nargs=argparse.REMAINDER, help="command to run and monitor (after --)")
    p.add_argument("-i", "--interval", type=float, default=1.0)
    p.add_argument("-o", "--out", "--output", dest="out", type=Path, default=Path("logs/mem_usage.csv"),
                   help="CSV with the samples (the reference's -o/--output)")
    p.add_argument("-p", "--plot", action="store_true", help="also write <out>.png")
    p.add_argument("--img", type=Path, default=None, help="plot to this image file (the reference's --img; implies --plot)")
    p.add_argument("-v", "--version", action="version", version="%(prog)s 0.2")
    a = p.parse_args(argv)
    if a.img is not None:
        a.plot = True
    cmd = [c for c in a.command if c != "--"]
    if not cmd:
        p.error("no command given")
    a.out.parent.mkdir(parents=True, exist_ok=True)
    proc = subprocess.Popen(cmd if len(cmd) > 1 else shlex.split(cmd[0]))
    n_gpu = len(gpu_used_mib())
    rows = []
    t0 = time.time()
    with open(a.out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["time_s", "rss_mib"] + [f"gpu{i}_used_mib" for i in range(n_gpu)])
        while proc.poll() is None:
            row = [round(time.time() - t0, 3), round(tree_rss_mib(proc), 1)] + [round(x, 1) for x in gpu_used_mib()]
            w.writerow(row)
            f.flush()
            rows.append(row)
            time.sleep(a.interval)
    print(f"command exited with {proc.returncode}; {len(rows)} samples in {a.out}")
    if rows:
        print(f"peak RSS {max(r[1] for r in rows):.0f} MiB" + (f", peak GPU0 {max(r[2] for r in rows):.0f} MiB" if n_gpu else ""))
    if a.plot:
        from ..utils.plots import have_matplotlib

        if have_matplotlib() and rows:
            import matplotlib

            matplotlib.use("Agg")
            import matplotlib.pyplot as plt

            plt.figure(figsize=(10, 5))
            plt.plot([r[0] for r in rows], [r[1] for r in rows], label="RSS (MiB)")
            for g in range(n_gpu):
                plt.plot([r[0] for r in rows], [r[2 + g] for r in rows], label=f"GPU{g} (MiB)")
            plt.xlabel("time (s)"); plt.legend(); plt.grid()
            img = a.img if a.img is not None else a.out.with_suffix(".png")
            img.parent.mkdir(parents=True, exist_ok=True)
            plt.savefig(img)
    return proc.returncode or 0


if __name__ == "__main__":
    raise SystemExit(main())
