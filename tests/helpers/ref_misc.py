"""Run in a subprocess by tests/test_reference_parity.py: small behaviours of the REFERENCE that are easy to miss —
stop-sequence truncation / detection on generated cases, and the prompt list built from a literal or a FILE: argument."""
import importlib
import json
import random
import sys

import torch

ref_root, shims, prompt_file, out_file = sys.argv[1:5]
sys.path.insert(0, shims)
sys.path.insert(0, ref_root)
import sub  # noqa: E402,F401
from sub.prompts import NoPrompt, PromptStyle, get_user_prompt  # noqa: E402

utils = importlib.import_module("sub.utils.utils")
rng = random.Random(7)
cases = []
for _ in range(300):
    n = rng.randint(2, 14)  # (a single token makes the reference's `.squeeze().tolist()` return an int)
    toks = [rng.randint(0, 5) for _ in range(n)]
    stops = tuple([rng.randint(0, 5) for _ in range(rng.randint(1, 3))] for _ in range(rng.randint(1, 3)))
    plen = rng.randint(0, n)
    t = torch.tensor([toks])
    try:
        cut = utils.find_eot(t, stops, plen).view(-1).tolist()
    except Exception as e:  # noqa: BLE001
        cut = f"ERR {type(e).__name__}"
    cases.append({"tokens": toks, "stops": [list(s) for s in stops], "prompt_length": plen, "find_eot": cut,
                  "detect": bool(utils.detect_stop_tokens(t, stops))})
prompts = {}
for key, (arg, n) in {"literal": ("Once upon a time", 3), "file_fewer": (f"FILE:{prompt_file}", 2), "file_more": (f"FILE:{prompt_file}", 5)}.items():
    for style_name in ("none", "alpaca"):
        style = NoPrompt() if style_name == "none" else PromptStyle.from_name("alpaca")
        try:
            prompts[f"{key}/{style_name}"] = get_user_prompt(arg, n, style)
        except Exception as e:  # noqa: BLE001
            prompts[f"{key}/{style_name}"] = f"ERR {type(e).__name__}"
# training helpers: learning-rate schedule and random batch extraction (same RNG consumption -> same batches)
lrs = [utils.get_lr(it, lr=3e-4, min_lr=3e-5, warmup_it=10, lr_decay_it=90) for it in range(0, 120, 3)]
dl = importlib.import_module("sub.utils.data_loader")


class _Conf:
    block_size = 8


data = torch.arange(1000) % 97
torch.manual_seed(5)
batches = []
for _ in range(3):
    x, y = dl.get_batch(data, 4, "cpu", _Conf())
    batches.append([x.tolist(), y.tolist()])
tr, va = dl.split_dataset(data, 0.9)
# the sampler: same logits + same RNG state -> same token (top-k crop, temperature, top-p, one multinomial draw)
from sub.model import sample  # noqa: E402

g = torch.Generator().manual_seed(99)
draws = []
for i in range(40):
    logits = torch.randn(1, 3, 50, generator=g) * 3
    kw = [dict(temperature=0.8, top_k=20), dict(temperature=1.0, top_k=None), dict(temperature=0.0, top_k=5, top_p=0.0),
          dict(temperature=0.7, top_k=200, top_p=0.9)][i % 4]
    torch.manual_seed(1000 + i)
    draws.append(int(sample(logits, **kw)))
json.dump({"cases": cases, "prompts": prompts, "lrs": lrs, "batches": batches, "split": [len(tr), len(va)], "draws": draws}, open(out_file, "w"))
