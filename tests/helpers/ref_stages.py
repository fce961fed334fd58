"""Run in a subprocess by tests/test_reference_parity.py: OUR chunk dictionaries loaded into the REFERENCE's stage
classes (``sub.submodels.StarterNode`` / ``SecondaryNode``), one pass of a prompt round the ring, logits out."""
import sys

import torch

ref_root, shims, cfg_file, chunks_file, out_file = sys.argv[1:6]
sys.path.insert(0, shims)
sys.path.insert(0, ref_root)
from sub.model import Config  # noqa: E402
from sub.submodels import SecondaryNode, StarterNode  # noqa: E402

cfg = Config(**torch.load(cfg_file))
chunks = torch.load(chunks_file)


def n_blocks(sd):
    return len({k.split(".")[2] for k in sd if k.startswith("transformer.h.")})


with torch.device("meta"):  # like gptserver.py:657-664: the weights replace meta tensors, nothing is allocated twice
    starter = StarterNode(cfg, n_blocks(chunks[0]))
starter.load_weights(chunks[0])
starter.cos, starter.sin = starter.rope_cache(device="cpu")  # the RoPE tables were built on meta with the module
secondaries = []
for sd in chunks[1:]:
    with torch.device("meta"):
        s = SecondaryNode(cfg, n_blocks(sd))
    s.load_weights(sd)
    s.cos, s.sin = s.rope_cache(device="cpu")
    secondaries.append(s.eval())
starter.eval()
idx = torch.tensor([[5, 17, 3, 44, 42, 7]])
with torch.no_grad():
    x = starter(idx)
    hidden = [x]
    for s in secondaries:
        x = s(x)
        hidden.append(x)
    logits = starter(x, first_pass=False)
torch.save({"hidden": hidden, "logits": logits}, out_file)
