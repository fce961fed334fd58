"""Run in a subprocess by tests/test_reference_parity.py: load a state dict into the REFERENCE's own model class
(``sub.model.GPT`` from the unmodified reference tree) and write its logits — teacher-forced, then cached decoding."""
import sys

import torch

ref_root, shims, cfg_file, sd_file, out_file = sys.argv[1:6]
sys.path.insert(0, shims)      # cherrypy / accelerate / matplotlib stand-ins: `import sub` pulls the whole package in
sys.path.insert(0, ref_root)
from sub.model import GPT, Config  # noqa: E402

kw = torch.load(cfg_file)
cfg = Config(**kw)
model = GPT(cfg)
sd = torch.load(sd_file)
model.load_state_dict(sd)
model.eval()
idx = torch.tensor([[5, 17, 3, 88, 42, 7]])
with torch.no_grad():
    full = model(idx)
    model.max_seq_length = 32
    model.set_kv_cache(batch_size=1)
    logits = model(idx, torch.arange(idx.size(1)))
    outs, toks = [logits[:, -1]], []
    for i in range(4):
        t = logits[:, -1].argmax(-1, keepdim=True)
        toks.append(int(t))
        logits = model(t, torch.tensor([idx.size(1) + i]))
        outs.append(logits[:, -1])
torch.save({"full": full, "dec": torch.stack(outs), "toks": toks}, out_file)
