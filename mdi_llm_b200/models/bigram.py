"""Toy language models of the reference's first generation (``old/nanoGPT/old_models/bigram.py``,
``bigram_attention.py``): a bigram table and a one-block attention model over characters.  They
exist for smoke-testing the data loader / trainer / tokenizers in seconds on a CPU.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

__all__ = ["BigramLanguageModel", "TinyAttentionLM"]


class _Generative(nn.Module):
    block_size: int = 1 << 30

    @torch.no_grad()
    def generate(self, idx: torch.Tensor, max_new_tokens: int, temperature: float = 1.0,
                 generator: Optional[torch.Generator] = None) -> torch.Tensor:
        for _ in range(max_new_tokens):
            logits, _ = self(idx[:, -self.block_size:])
            logits = logits[:, -1]
            if temperature == 0.0:
                nxt = logits.argmax(-1, keepdim=True)
            else:
                nxt = torch.multinomial(F.softmax(logits / temperature, dim=-1), 1, generator=generator)
            idx = torch.cat((idx, nxt), dim=1)
        return idx


class BigramLanguageModel(_Generative):
    """Next-token logits are one row of a ``V x V`` table."""

    def __init__(self, vocab_size: int) -> None:
        super().__init__()
        self.table = nn.Embedding(vocab_size, vocab_size)

    def forward(self, idx: torch.Tensor, targets: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        logits = self.table(idx)
        loss = None if targets is None else F.cross_entropy(logits.view(-1, logits.size(-1)), targets.reshape(-1))
        return logits, loss


class TinyAttentionLM(_Generative):
    """Token + position embeddings, ``n_layer`` pre-norm attention/MLP blocks, linear head."""

    def __init__(self, vocab_size: int, n_embd: int = 32, n_head: int = 4, n_layer: int = 1, block_size: int = 32) -> None:
        super().__init__()
        self.block_size = block_size
        self.tok = nn.Embedding(vocab_size, n_embd)
        self.pos = nn.Embedding(block_size, n_embd)
        self.blocks = nn.ModuleList(
            nn.TransformerEncoderLayer(n_embd, n_head, 4 * n_embd, dropout=0.0, batch_first=True, norm_first=True)
            for _ in range(n_layer))
        self.ln = nn.LayerNorm(n_embd)
        self.head = nn.Linear(n_embd, vocab_size)

    def forward(self, idx: torch.Tensor, targets: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        T = idx.size(1)
        x = self.tok(idx) + self.pos(torch.arange(T, device=idx.device))
        mask = torch.triu(torch.full((T, T), float("-inf"), device=idx.device), diagonal=1)
        for b in self.blocks:
            x = b(x, src_mask=mask)
        logits = self.head(self.ln(x))
        loss = None if targets is None else F.cross_entropy(logits.view(-1, logits.size(-1)), targets.reshape(-1))
        return logits, loss
