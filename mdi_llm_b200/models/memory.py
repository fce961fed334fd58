"""HBM budget of a pipeline stage: does a plan fit the 180 GB of a B200 before anything is allocated?

The partition planners balance *time* (``models/partition.py``); a big model on few GPUs is first of all a *memory*
problem (Falcon-180B in bf16 is 360 GB of weights, Mixtral-8x7B keeps all 8 experts of every layer resident although a
token streams 2).  ``stage_memory`` prices one stage exactly the way the runtime allocates it:

* weights — counted on a meta-device instance of the very stage module that will be built (no formula per model
  family to drift), 2 bytes per parameter, or 1 byte + one fp32 scale per 128 for the projections with ``weights="fp8"``
  (embeddings, norms and biases stay bf16, ``parallel/engine.py::_quantize``);
* KV slots — ``KVPool``: ``[local blocks, n_samples, 2, G, S, head_size]`` bf16 (G group heads, not H);
* hop buffers — the decode row and the prefill buffer of every sample slot (``[n_samples, max_prompt_len, W_in]``),
  plus on the starter the fp32 logits row and the token ring.

``check_plan`` is what ``GPTDistributed`` calls on a CUDA starter before it sends ``/init``: a stage that cannot fit is
reported with its numbers instead of an out-of-memory error half-way through loading a 100 GB chunk.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

from .config import Config

B200_HBM_BYTES = 180 * 10 ** 9  # what the runtime can count on per GPU (the device reports slightly more)
_PROJ_TAILS = (".attn.attn.weight", ".attn.proj.weight", ".mlp.fc.weight", ".mlp.fc_1.weight", ".mlp.fc_2.weight", ".mlp.proj.weight")


def stage_memory(config: Config, spec: Dict[str, Any], is_starter: bool, n_samples: int, max_seq_length: int,
                 max_prompt_len: Optional[int] = None, weights: str = "bf16") -> Dict[str, int]:
    """Bytes a stage allocates: ``{"weights", "kv", "hop", "total"}``.  ``spec``: an entry of ``stage_specs``."""
    from .stage import build_stage

    kw = {k: spec[k] for k in ("first_parts", "last_parts") if spec.get(k)}
    st = build_stage(config, "starter" if is_starter else "secondary:0", spec["n_blocks"], meta=True, **kw)
    w = 0
    for name, p in st.named_parameters():
        n = p.numel()
        quantised = weights == "fp8" and (name.endswith(_PROJ_TAILS) or (name == "lm_head.weight" and not config.tie_embeddings)
                                         or (".mlp.experts." in name and name.endswith(".weight")))
        w += n + (n // 128) * 4 if quantised else 2 * n
    if config.tie_embeddings and is_starter and any(n == "lm_head.weight" for n, _ in st.named_parameters()):
        w -= 2 * config.padded_vocab_size * config.n_embd  # one tensor seen under two names
    kv = spec["n_blocks"] * n_samples * 2 * config.n_query_groups * max_seq_length * config.head_size * 2
    w_in = int(getattr(st, "in_width", config.n_embd))
    prompt = int(max_prompt_len or max_seq_length)
    hop = n_samples * w_in * 2 + n_samples * prompt * w_in * 2
    if is_starter:
        hop += config.padded_vocab_size * 4 + n_samples * (max_seq_length + 1) * (4 + 8)
    return {"weights": int(w), "kv": int(kv), "hop": int(hop), "total": int(w + kv + hop)}


def plan_memory(config: Config, specs: Sequence[Dict[str, Any]], n_samples: int, max_seq_length: int,
                max_prompt_len: Optional[int] = None, weights: str = "bf16") -> List[Dict[str, int]]:
    return [stage_memory(config, sp, i == 0, n_samples, max_seq_length, max_prompt_len, weights) for i, sp in enumerate(specs)]


def check_plan(config: Config, specs: Sequence[Dict[str, Any]], n_samples: int, max_seq_length: int,
               max_prompt_len: Optional[int] = None, weights: str = "bf16", capacity: int = B200_HBM_BYTES,
               headroom: float = 0.94) -> List[str]:
    """One message per stage that does not fit ``headroom * capacity`` (the rest is the CUDA context, the allocator's
    slack and the prompt-sized temporaries of the prefill GEMMs); empty list = the plan fits."""
    out = []
    for i, m in enumerate(plan_memory(config, specs, n_samples, max_seq_length, max_prompt_len, weights)):
        if m["total"] > headroom * capacity:
            out.append(f"stage {i} ({specs[i].get('layers', specs[i]['n_blocks'])} layers) needs {m['total'] / 1e9:.1f} GB "
                       f"(weights {m['weights'] / 1e9:.1f}, KV {m['kv'] / 1e9:.1f} for {n_samples} samples x {max_seq_length} positions, "
                       f"buffers {m['hop'] / 1e9:.2f}) of {capacity / 1e9:.0f} GB: use more nodes, `--weights fp8`, fewer samples "
                       f"or a shorter --sequence-length")
    return out
