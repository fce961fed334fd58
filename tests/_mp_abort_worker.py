"""Abort propagation: 3-node device ring through the node API; the last secondary is frozen (SIGSTOP: it stops
enqueuing steps, its memory stays mapped) after a few healthy rounds.  The starter must come back with an error
within a few watchdog periods (not hang), and secondary 0 must report `aborted` (poison reached it)."""
import json
import os
import signal
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from mdi_llm_b200.cli.launch import loopback_topology  # noqa: E402
from mdi_llm_b200.parallel.control import call_node  # noqa: E402
from mdi_llm_b200.parallel.distributed import GPTDistributed  # noqa: E402
from mdi_llm_b200.parallel.ring import RingError  # noqa: E402
from mdi_llm_b200.parallel.scheduler import SamplingParams  # noqa: E402


def main():
    ck = sys.argv[1]
    n = min(3, torch.cuda.device_count())
    topo = loopback_topology(n)
    topo_file = os.path.join(os.path.dirname(ck), "abort_nodes.json")
    with open(topo_file, "w") as f:
        json.dump(topo, f)
    env = dict(os.environ, PYTHONPATH=ROOT)
    secs = [subprocess.Popen([sys.executable, "-m", "mdi_llm_b200.cli.secondary", "--nodes-config", topo_file, str(i), "--ckpt", ck],
                             env=env) for i in range(n - 1)]
    try:
        gd = GPTDistributed("starter", topo, ckpt_dir=ck, model_seq_length=128, sampling=SamplingParams.greedy(), watchdog_s=1.0)
        sess = gd.open_session(n + 1, 40, "Hello there")
        sess.run(4)  # healthy so far
        victim = secs[-1]
        victim.send_signal(signal.SIGSTOP)
        sess.timeout = 8.0  # control-plane requests to the frozen node give up quickly
        t0 = time.time()
        try:
            sess.run(20)
            print("ABORT_RESULT no error raised")
            return 1
        except RingError as e:
            dt = time.time() - t0
            ok = dt < 30
            detail = {"seconds": round(dt, 2), "error": str(e)[:200]}
            if n == 3:  # the surviving secondary saw the poison (or tripped itself) and drained
                s0 = topo["nodes"]["secondary"][0]
                status, body = call_node("post", f"http://{s0['addr']}:{s0['communication']['port']}/ring", {"op": "stats"},
                                         max_n_requests=1)
                detail["sec0"] = body
                ok = ok and status == 200 and any(body["status"])
            print("ABORT_RESULT " + ("ok " if ok else "bad ") + json.dumps(detail))
        finally:
            victim.send_signal(signal.SIGCONT)
            sess.close()
            gd.stop_nodes()
            gd.gpt_serv.shutdown()
    finally:
        for p in secs:
            if p.poll() is None:
                p.terminate()
        for p in secs:
            try:
                p.wait(timeout=10)
            except subprocess.TimeoutExpired:
                p.kill()
    return 0


if __name__ == "__main__":
    sys.exit(main())
