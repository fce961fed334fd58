"""Run in a subprocess by tests/test_reference_parity.py: the REFERENCE's own ``split_parameters`` on a state dict."""
import importlib
import sys

import torch

ref_root, shims, sd_file, n_nodes, out_file = sys.argv[1:6]
sys.path.insert(0, shims)
sys.path.insert(0, ref_root)
import sub  # noqa: E402,F401

utils = importlib.import_module("sub.utils.utils")
chunks, info = utils.split_parameters(torch.load(sd_file), int(n_nodes))
torch.save({"chunks": chunks, "info": dict(info)}, out_file)
