#!/usr/bin/env python3
"""One-box launcher: every node of a model-distributed run as a process of THIS machine.

Parity: the reference's ``old/nanoGPT/test_mdi_local.sh:1-53`` (spawn finisher / intermediate / starter in
the background, N runs, kill stragglers) and ``test_local_gen.sh`` (repeated timed runs), plus SURVEY §7.1's
"in-box launcher (spawn one proc per GPU)".  Here it is one command:

    python -m mdi_llm_b200.cli.launch --ckpt <dir> --n-nodes 4 --runs 3 -- --n-samples 4 --n-tokens 200 -p

* without ``--nodes-config`` a loopback topology is generated (free ports, node *i* on ``cuda:i`` — or on the
  CPU with ``--device cpu``), in the reference's JSON schema;
* secondaries are started first (they wait for ``POST /init``), then the starter runs in the foreground with
  everything after ``--`` passed through (``--time-run`` collects one CSV row per run);
* stragglers are terminated **by the PIDs this launcher started** — never by name pattern.
"""
from __future__ import annotations

import argparse
import json
import os
import signal
import socket
import subprocess
import sys
import tempfile
import time
from pathlib import Path
from typing import Any, Dict, List, Optional


def free_ports(n: int) -> List[int]:
    socks, ports = [], []
    for _ in range(n):
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        socks.append(s)
        ports.append(s.getsockname()[1])
    for s in socks:
        s.close()
    return ports


def loopback_topology(n_nodes: int, device: str = "cuda") -> Dict[str, Any]:
    """Node JSON in the reference's schema (src/settings_distr/*.json) for ``n_nodes`` processes on 127.0.0.1."""
    ports = free_ports(3 * n_nodes)

    def node(i: int) -> Dict[str, Any]:
        dev = f"cuda:{i}" if device == "cuda" else device
        return {"addr": "127.0.0.1", "communication": {"port": ports[3 * i], "starter_addr": "127.0.0.1"},
                "inference": {"port_in": ports[3 * i + 1], "port_out": ports[3 * i + 2]}, "device": dev}

    return {"nodes": {"starter": node(0), "secondary": [node(i) for i in range(1, n_nodes)]}}


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Launch every node of an MDI run on this machine",
                                epilog="arguments after `--` go to the starter CLI unchanged")
    p.add_argument("--ckpt", type=Path, required=True, help="checkpoint folder (lit_model.pth / model_config.yaml / chunks)")
    p.add_argument("--n-nodes", type=int, default=None, help="number of nodes when no --nodes-config is given")
    p.add_argument("--nodes-config", type=Path, default=None, help="existing topology JSON (else a loopback one is generated)")
    p.add_argument("--device", default="cuda", help="cuda (node i on cuda:i) | cpu | an explicit torch device for every node")
    p.add_argument("--dtype", default=None)
    p.add_argument("--runs", type=int, default=1, help="repeat the whole launch this many times (test_mdi_local.sh)")
    p.add_argument("--secondary-grace", type=float, default=15.0, help="seconds to wait for secondaries to exit after a run")
    p.add_argument("-v", "--verb", action="store_true")
    p.add_argument("starter_args", nargs=argparse.REMAINDER, help="-- <flags of mdi_llm_b200.cli.starter>")
    return p


def _stop(procs: List[subprocess.Popen], grace: float) -> None:
    deadline = time.time() + grace
    for pr in procs:  # PUT /stop from the starter normally ends them
        try:
            pr.wait(timeout=max(0.0, deadline - time.time()))
        except subprocess.TimeoutExpired:
            pass
    for pr in procs:
        if pr.poll() is None:
            pr.send_signal(signal.SIGTERM)
    for pr in procs:
        try:
            pr.wait(timeout=5)
        except subprocess.TimeoutExpired:
            pr.kill()


def run_once(args: argparse.Namespace, topo_file: Path, n_nodes: int, extra: List[str]) -> int:
    env = dict(os.environ)
    root = str(Path(__file__).resolve().parents[2])
    env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
    common = ["--ckpt", str(args.ckpt)] + (["--dtype", args.dtype] if args.dtype else []) + (["-v"] if args.verb else [])
    secs: List[subprocess.Popen] = []
    try:
        for i in range(n_nodes - 1):
            cmd = [sys.executable, "-m", "mdi_llm_b200.cli.secondary", "--nodes-config", str(topo_file), str(i)] + common
            secs.append(subprocess.Popen(cmd, env=env))
        cmd = [sys.executable, "-m", "mdi_llm_b200.cli.starter", "--nodes-config", str(topo_file)] + common + extra
        rc = subprocess.call(cmd, env=env)
    finally:
        _stop(secs, args.secondary_grace)
    return rc


def main(argv: Optional[List[str]] = None) -> int:
    args = build_parser().parse_args(argv)
    extra = [a for a in args.starter_args if a != "--"] if args.starter_args[:1] == ["--"] else list(args.starter_args)
    tmp: Optional[tempfile.TemporaryDirectory] = None
    rc = 0
    try:
        for run in range(args.runs):
            if args.nodes_config is not None:
                topo_file = args.nodes_config
                with open(topo_file) as f:
                    n_nodes = 1 + len(json.load(f)["nodes"].get("secondary", []))
            else:
                if not args.n_nodes or args.n_nodes < 1:
                    raise SystemExit("--n-nodes or --nodes-config is required")
                tmp = tmp or tempfile.TemporaryDirectory(prefix="mdi_launch_")
                topo_file = Path(tmp.name) / f"nodes_{run}.json"  # fresh ports every run (TIME_WAIT)
                topo_file.write_text(json.dumps(loopback_topology(args.n_nodes, args.device), indent=1))
                n_nodes = args.n_nodes
            print(f"=== run {run + 1}/{args.runs}: {n_nodes} node(s), topology {topo_file} ===", flush=True)
            rc = run_once(args, topo_file, n_nodes, extra)
            if rc != 0:
                print(f"run {run + 1} failed with exit code {rc}", file=sys.stderr)
                break
    finally:
        if tmp is not None:
            tmp.cleanup()
    return rc


if __name__ == "__main__":
    raise SystemExit(main())
