"""Run in a subprocess by tests/test_reference_parity.py: one node of the UNMODIFIED reference (its own
``GPTDistributed``), as starter or as secondary, on CPU."""
import sys

import torch

ref_root, shims, role, topo_file, ckpt_dir = sys.argv[1:6]
sys.path.insert(0, shims)
sys.path.insert(0, ref_root)
from sub.model_dist import GPTDistributed  # noqa: E402

torch.manual_seed(1337)
if role == "starter":
    n_samples, n_tokens, prompt = int(sys.argv[6]), int(sys.argv[7]), sys.argv[8]
    # model_seq_length: the reference rebuilds its RoPE tables only when the context is truncated (submodels.py:50-66); with
    # the full block size its tables stay whatever `.to(device)` makes of meta tensors
    node = GPTDistributed("starter", topo_file, ckpt_dir=ckpt_dir, device="cpu", dtype="float32", verb=False, plots=True,  # (without plots the reference indexes an empty timeline at the end)
                          model_seq_length=int(sys.argv[9]) if len(sys.argv) > 9 else None)
    node.start(n_samples=n_samples, tokens_per_sample=n_tokens, prompt=prompt)
else:
    # "-" = no checkpoint on this node's disk: it is model-agnostic until POST /init brings config and chunk
    node = GPTDistributed(role, topo_file, ckpt_dir=None if ckpt_dir == "-" else ckpt_dir, device="cpu", dtype="float32", verb=False)
    node.start()  # blocks until the starter's PUT /stop
