"""Run in a subprocess by tests/test_reference_parity.py: the FIRST-GENERATION GPT-2 model of the reference
(``old/GPT2/sub/model.py``, nanoGPT layout) with random weights: its state dict and its logits for a prompt.
(The module is loaded without the package's ``__init__`` — that one imports the whole distributed stack.)"""
import importlib.util
import sys
import types

import torch

old_sub, out_file = sys.argv[1:3]
pkg = types.ModuleType("oldsub")
pkg.__path__ = [old_sub]
sys.modules["oldsub"] = pkg


def load(name):
    spec = importlib.util.spec_from_file_location(f"oldsub.{name}", f"{old_sub}/{name}.py")
    m = importlib.util.module_from_spec(spec)
    sys.modules[f"oldsub.{name}"] = m
    spec.loader.exec_module(m)
    return m


load("config")
model = load("model")
torch.manual_seed(3)
cfg = model.GPTConfig(block_size=32, vocab_size=128, n_layer=2, n_head=4, n_embd=64, dropout=0.0, bias=True)
m = model.GPT(cfg).eval()
with torch.no_grad():
    for p in m.parameters():  # nanoGPT initialises biases to zero: make every tensor count
        p.add_(torch.randn_like(p) * 0.05)
idx = torch.tensor([[5, 17, 3, 88, 42, 7]])
with torch.no_grad():
    rows = [m(idx[:, : t + 1])[0][:, -1] for t in range(idx.size(1))]  # inference returns the last position only
torch.save({"sd": {k: v.clone() for k, v in m.state_dict().items()}, "logits": torch.stack(rows, dim=1)}, out_file)
