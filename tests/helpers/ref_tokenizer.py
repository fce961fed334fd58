"""Run in a subprocess by tests/test_reference_parity.py: the REFERENCE's Tokenizer on a checkpoint directory."""
import json
import sys

ref_root, shims, ckpt_dir, out_file = sys.argv[1:5]
sys.path.insert(0, shims)
sys.path.insert(0, ref_root)
from sub.tokenizer import Tokenizer  # noqa: E402

TEXTS = ["Shall I compare thee to a summer's day?", "  leading spaces and a\nnewline", "", "Thou art more lovely — and more temperate: 123!"]
tok = Tokenizer(ckpt_dir)
out = {"backend": tok.backend, "bos_id": tok.bos_id, "eos_id": tok.eos_id, "use_bos": tok.use_bos, "vocab_size": tok.vocab_size, "cases": []}
for t in TEXTS:
    for bos, eos, max_len in ((None, False, -1), (True, True, -1), (False, True, 5)):
        try:
            ids = tok.encode(t, bos=bos, eos=eos, max_length=max_len).tolist()
            out["cases"].append({"text": t, "bos": bos, "eos": eos, "max_length": max_len, "ids": ids, "decoded": tok.decode(tok.encode(t, bos=False))})
        except Exception as e:  # noqa: BLE001
            out["cases"].append({"text": t, "bos": bos, "eos": eos, "max_length": max_len, "error": type(e).__name__})
json.dump(out, open(out_file, "w"))
