"""Run in a subprocess by tests/test_reference_parity.py: the REFERENCE's own weight-conversion functions, both ways.
lit state dict -> HF names (convert_lit_checkpoint.copy_weights_*), then that HF dict -> lit (convert_hf_checkpoint.*)."""
import sys
from functools import partial

import torch

ref_root, shims, cfg_file, sd_file, out_file = sys.argv[1:6]
sys.path.insert(0, shims)
sys.path.insert(0, ref_root)
import importlib  # noqa: E402

from sub.model import Config  # noqa: E402

# (`sub.utils` re-exports functions under the modules' own names: fetch the modules themselves)
to_lit = importlib.import_module("sub.utils.convert_hf_checkpoint")
to_hf = importlib.import_module("sub.utils.convert_lit_checkpoint")

config = Config(**torch.load(cfg_file))
lit = torch.load(sd_file)

if "falcon" in config.name:
    fn = partial(to_hf.copy_weights_falcon, config.name)
elif config.mlp_class_name in ("LLaMAMLP", "GemmaMLP", "LLaMAMoE"):
    fn = partial(to_hf.copy_weights_llama, config, untie_weights="Gemma" in config.name)
elif "phi" in config.name:
    fn = partial(to_hf.copy_weights_phi, config)
else:
    fn = to_hf.copy_weights_gpt_neox
hf = {}
fn(hf, dict(lit))

if "falcon" in config.name:
    back_fn = partial(to_lit.copy_weights_falcon, config.name)
elif config.mlp_class_name in ("LLaMAMLP", "GemmaMLP", "LLaMAMoE"):
    back_fn = partial(to_lit.copy_weights_hf_llama, config, {})
elif "phi" in config.name:
    back_fn = partial(to_lit.copy_weights_phi, config, {})
else:
    back_fn = to_lit.copy_weights_gpt_neox
back, err = {}, None
try:
    back_fn(back, dict(hf))
except NameError as e:  # the reference's Phi import path uses `defaultdict` without importing it (convert_hf_checkpoint.py:241)
    back, err = None, repr(e)
torch.save({"hf": hf, "lit": back, "hf_to_lit_error": err}, out_file)
