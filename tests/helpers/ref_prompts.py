"""Run in a subprocess by tests/test_reference_parity.py: what the REFERENCE's prompt module produces — every style
applied to two prompts, and the style its regex dispatch picks for every registry model name."""
import json
import sys

ref_root, shims, out_file = sys.argv[1:4]
sys.path.insert(0, shims)
sys.path.insert(0, ref_root)
from sub.config import configs  # noqa: E402
from sub.prompts import PromptStyle, model_name_to_prompt_style, prompt_styles  # noqa: E402

PROMPTS = ["Hello, how are you?", "Write a haiku about GPUs.\nMake it rhyme."]
applied = {}
for name in prompt_styles:
    style = PromptStyle.from_name(name)
    try:
        applied[name] = [style.apply(p) for p in PROMPTS]
    except Exception as e:  # noqa: BLE001
        applied[name] = f"ERR {type(e).__name__}"
names = sorted({c["name"] for c in configs} | {c["hf_config"]["name"] for c in configs})
picked = {n: type(model_name_to_prompt_style(n)).__name__ for n in names}
json.dump({"applied": applied, "picked": picked}, open(out_file, "w"))
