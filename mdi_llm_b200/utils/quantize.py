"""fp8 (e4m3) block-scaled weight quantisation for serving.

New capability relative to the reference (its dtypes stop at fp32/bf16/fp16, ``config.py:109-114``;
BASELINE.json config #5 asks for "Llama-3-8B fp8 block-scaled").  Scheme: every weight row is cut
into blocks of 128 consecutive input channels; a block stores e4m3 values and one fp32 scale
``amax / 448``.  Decode kernels consume this directly (``stream_ldg_fp8_kernel``: bytes per token
halved); prefill dequantises one matrix at a time into a bf16 scratch and runs the tcgen05 GEMM.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

__all__ = ["FP8_BLOCK", "quantize_fp8_block", "dequantize_fp8_block", "quantize_linear_weights", "fp8_error"]

FP8_BLOCK = 128
_E4M3_MAX = 448.0


def quantize_fp8_block(w: torch.Tensor, block: int = FP8_BLOCK) -> Tuple[torch.Tensor, torch.Tensor]:
    """``w [N, K]`` -> (``q [N, K]`` float8_e4m3fn, ``scale [N, K/block]`` fp32) with ``w ≈ q * scale``."""
    if w.dim() != 2 or w.shape[1] % block:
        raise ValueError(f"weight of shape {tuple(w.shape)}: K must be a multiple of {block}")
    n, k = w.shape
    wf = w.float().view(n, k // block, block)
    scale = (wf.abs().amax(dim=-1) / _E4M3_MAX).clamp_min(1e-12)
    q = (wf / scale.unsqueeze(-1)).clamp(-_E4M3_MAX, _E4M3_MAX).to(torch.float8_e4m3fn).view(n, k)
    return q.contiguous(), scale.contiguous()


def dequantize_fp8_block(q: torch.Tensor, scale: torch.Tensor, dtype: torch.dtype = torch.bfloat16,
                         block: int = FP8_BLOCK) -> torch.Tensor:
    n, k = q.shape
    return (q.float().view(n, k // block, block) * scale.unsqueeze(-1)).view(n, k).to(dtype)


def quantize_linear_weights(state_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Quantise every 2-D projection weight of a litGPT state dict (embeddings and norms stay in
    their dtype).  ``<name>.weight`` becomes fp8 and ``<name>.weight_scale`` is added."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in state_dict.items():
        is_proj = k.endswith(".weight") and v.dim() == 2 and "wte" not in k and "wpe" not in k and v.shape[1] % FP8_BLOCK == 0
        if is_proj:
            q, s = quantize_fp8_block(v)
            out[k], out[k + "_scale"] = q, s
        else:
            out[k] = v
    return out


def fp8_error(w: torch.Tensor) -> float:
    """Relative Frobenius error of the round trip (accuracy gate used by the tests)."""
    q, s = quantize_fp8_block(w)
    d = dequantize_fp8_block(q, s, torch.float32)
    return float((d - w.float()).norm() / w.float().norm().clamp_min(1e-12))
