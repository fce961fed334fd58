"""Plain-PyTorch scaled-dot-product attention.

Parity: reference ``src/sub/utils/functional.py:7-33`` (a fallback for torch < 2.0 on Jetson).
Here it serves as the fp32 *oracle* for the CUDA attention kernels' numerics tests and keeps the
reference's ``__main__`` self-check (:36-74) against ``F.scaled_dot_product_attention``.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

__all__ = ["scaled_dot_product_attention"]


def scaled_dot_product_attention(
    query: torch.Tensor,
    key: torch.Tensor,
    value: torch.Tensor,
    attn_mask: Optional[torch.Tensor] = None,
    dropout_p: float = 0.0,
    is_causal: bool = False,
    scale: Optional[float] = None,
) -> torch.Tensor:
    L, S = query.size(-2), key.size(-2)
    scale = 1.0 / math.sqrt(query.size(-1)) if scale is None else scale
    scores = (query @ key.transpose(-2, -1)) * scale
    if is_causal:
        if attn_mask is not None:
            raise ValueError("pass either attn_mask or is_causal")
        keep = torch.ones(L, S, dtype=torch.bool, device=query.device).tril()
        scores = scores.masked_fill(~keep, float("-inf"))
    if attn_mask is not None:
        if attn_mask.dtype == torch.bool:
            scores = scores.masked_fill(~attn_mask, float("-inf"))
        else:
            scores = scores + attn_mask
    probs = torch.softmax(scores, dim=-1)
    if dropout_p > 0.0:
        probs = torch.dropout(probs, dropout_p, train=True)
    return probs @ value


if __name__ == "__main__":
    import torch.nn.functional as F

    torch.manual_seed(0)
    q, k, v = (torch.randn(1, 1, 10, 10) for _ in range(3))
    mask = torch.ones(10, 10, dtype=torch.bool).tril()[None, None]
    ref = F.scaled_dot_product_attention(q, k, v, attn_mask=mask)
    ours = scaled_dot_product_attention(q, k, v, attn_mask=mask)
    print("max |F.sdpa - python sdpa| =", (ref - ours).abs().max().item())
