"""CUDA stage executor: one pipeline stage's decode step as a short chain of fused sm_100a kernels.

Per transformer block the reference launches ≈35-40 ATen kernels from Python; here a
block is 6 launches (5 weight-streaming kernels + attention combine) and a whole stage step is
replayed as one CUDA graph::

    [hop wait]→ QKV(+RMSNorm+RoPE+KV append) → attention(split-KV, GQA packed) → out-proj(+residual)
              → gate/up(+RMSNorm+SiLU·mul) → down(+residual [+P2P store to next stage + flag])

The starter additionally runs ``lm_head(+final RMSNorm)`` → on-device sampling → embedding gather
at the start of its step.  Weights are used in place (the ``StageModule`` parameters), the KV
pool is shared with the eager module, which still serves prefill (T > 1) in this round.

Two ways to drive it:

* :class:`FusedStageRunner` — the :class:`~.scheduler.StageRunner` interface (host-driven; any
  transport).  Each call is: tiny H2D of the step descriptor → graph replay → result tensor.
* :class:`DevicePipeline` (``parallel/pipeline.py``) — device-driven ring over NVLink with fused
  hops; the host only enqueues graph replays.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Sequence, Any, Dict, List, Optional, Tuple

import torch

from .. import ops
from ..models.config import Config
from ..models.stage import StageModule, StarterNode
from .scheduler import SamplingParams, StageRunner

__all__ = ["engine_supports", "fused_prefill_supports", "FusedStage", "FusedStageRunner", "HopTarget"]


def engine_supports(config: Config, dtype: torch.dtype) -> bool:
    """Architectures the fused decode kernels cover (the rest runs on the eager runner): RMSNorm or LayerNorm,
    sequential or parallel residual (with or without a shared attention norm), gated (SwiGLU / GeGLU) or plain
    GELU MLPs or a routed mixture of SwiGLU experts, rotary (full or partial) or learned positions, head sizes
    64 / 128 / 256 — i.e. the Llama, Mistral, Mixtral, TinyLlama, Gemma, Pythia / GPT-NeoX, StableLM, GPT-2 families.
    Not covered: odd head sizes (Phi: 80) and very wide GQA groups (Falcon-7B: 71 query heads per KV head)."""
    moe = config.mlp_class_name == "LLaMAMoE"
    return (
        dtype == torch.bfloat16
        and config.norm_class_name in ("RMSNorm", "LayerNorm")
        and config.mlp_class_name in ("LLaMAMLP", "GemmaMLP", "GptNeoxMLP", "LLaMAMoE")
        and (not moe or (not config.parallel_residual and not config.bias and 1 <= config.n_expert_per_token <= 8
                         and config.n_expert_per_token <= config.n_expert <= 256))
        and config.pos_embedding in ("rope", "learned")
        and config.head_size in (64, 128, 256)
        and config.q_per_kv in (1, 2, 4, 8)
        and not (config.head_size == 256 and config.q_per_kv > 2)
        and config.n_embd % 8 == 0 and config.intermediate_size % 8 == 0
        and config.rope_n_elem % 2 == 0 and (config.rope_n_elem > 0 or config.pos_embedding == "learned")
        and (config.parallel_residual or not config.shared_attention_norm)
    )


def fused_prefill_supports(config: Config) -> bool:
    """Prompt processing on the tcgen05 GEMM + attention kernels (the Llama-shaped subset); other covered
    architectures prefill through their eager modules and decode through the fused kernels."""
    mlps = ("LLaMAMLP", "GemmaMLP") if os.environ.get("MDI_MOE_PREFILL", "gemm") == "eager" else ("LLaMAMLP", "GemmaMLP", "LLaMAMoE")
    return (config.norm_class_name == "RMSNorm" and not config.parallel_residual
            and config.mlp_class_name in mlps and config.pos_embedding == "rope" and config.rope_n_elem > 0
            and config.head_size in (64, 128))


@dataclass
class HopTarget:
    """Where the last block's epilogue stores the stage output: raw device pointers of the NEXT
    stage's ``hidden_in [n_slots, C]`` and ``flags [n_slots]`` (peer-mapped or local)."""

    hidden_ptr: int
    flag_ptr: int


class RawBuffer:
    """cudaMalloc'ed (IPC-exportable) memory with torch views on top."""

    def __init__(self, nbytes: int, device: torch.device) -> None:
        import ctypes

        self.nbytes = nbytes
        self.device = device
        p = ctypes.c_void_p()
        self._handle = ctypes.create_string_buffer(64)
        with torch.cuda.device(device):
            ops.check(ops.lib().mdi_p2p_alloc(nbytes, ctypes.byref(p), self._handle), "p2p_alloc")
        self.ptr = int(p.value)

    @property
    def handle(self) -> bytes:
        return bytes(self._handle.raw)

    def view(self, offset: int, shape: Tuple[int, ...], dtype: torch.dtype) -> torch.Tensor:
        n = int(torch.tensor(shape).prod().item()) if shape else 1
        typestr = {torch.bfloat16: "<u2", torch.int32: "<i4", torch.float32: "<f4", torch.uint8: "|u1"}[dtype]

        class _Iface:
            pass

        holder = _Iface()
        holder.__cuda_array_interface__ = {  # type: ignore[attr-defined]
            "shape": (n,), "typestr": typestr, "data": (self.ptr + offset, False), "version": 3,
        }
        t = torch.as_tensor(holder, device=self.device)
        if dtype == torch.bfloat16:
            t = t.view(torch.bfloat16)
        return t.view(*shape)

    def free(self) -> None:
        if self.ptr:
            ops.lib().mdi_p2p_free(self.ptr)
            self.ptr = 0


class FusedStage:
    """Buffers + kernel sequence of one stage.  All methods enqueue on the current stream."""

    def __init__(self, model: StageModule, n_slots: int, max_seq_length: Optional[int] = None,
                 sampling: Optional[SamplingParams] = None, use_pdl: bool = True, ctas_per_sm: int = 4,
                 wait_max_cycles: int = 0, exportable: bool = False, weight_dtype: str = "bf16",
                 free_bf16: bool = False) -> None:
        ops.require()
        cfg = model.config
        p = next(model.parameters())
        if not engine_supports(cfg, p.dtype):
            raise ValueError(f"config {cfg.name!r} / dtype {p.dtype} is not supported by the fused engine")
        self.model, self.cfg = model, cfg
        self.device = p.device
        self.is_starter = isinstance(model, StarterNode)
        self.n_layers = model.n_local_layers
        self.n_slots = n_slots
        self.S = int(max_seq_length or model.max_seq_length)
        if self.S != model.max_seq_length:
            model.max_seq_length = self.S
        self.sampling = sampling or SamplingParams()
        self.use_pdl, self.ctas_per_sm, self.wait_max_cycles = use_pdl, ctas_per_sm, wait_max_cycles
        C, dev = cfg.n_embd, self.device
        with torch.cuda.device(dev):
            model.cos, model.sin = model.cos.to(dev, torch.float32).contiguous(), model.sin.to(dev, torch.float32).contiguous()
            if model.kv_pool is None or model.kv_pool.n_slots < n_slots or model.kv_pool.max_seq_length != self.S:
                model.kv_pool = None
                model.set_kv_cache(n_slots, device=dev, dtype=torch.bfloat16)
            self.kv = model.kv_pool.data  # [L, n_slots, 2, G, S, hs]
            i32 = dict(dtype=torch.int32, device=dev)
            bf = dict(dtype=torch.bfloat16, device=dev)
            # message widths: C, or C + I when the stage starts at a down projection / ends after a gate/up unit
            # (third-layer plans, models/partition.py); hop-visible region: hidden_in [n_slots, W_in] bf16 | flags
            self.W_in = int(getattr(model, "in_width", C))
            self.W_out = int(getattr(model, "out_width", C))
            hid_bytes = n_slots * self.W_in * 2
            self._flag_off = (hid_bytes + 255) // 256 * 256
            self.raw = RawBuffer(self._flag_off + max(256, n_slots * 4), dev) if exportable else None
            if self.raw is not None:
                self.hidden_in = self.raw.view(0, (n_slots, self.W_in), torch.bfloat16)
                self.flags = self.raw.view(self._flag_off, (n_slots,), torch.int32)
            else:
                self.hidden_in = torch.zeros(n_slots, self.W_in, **bf)
                self.flags = torch.zeros(n_slots, **i32)
            self.out_local = torch.zeros(n_slots, self.W_out, **bf)  # the stage's output row (copied to the next stage by the hop)
            self.ctx = torch.zeros(ops.CTX_INTS, **i32)
            self.ctx_ring = torch.zeros(4096, ops.CTX_INTS, dtype=torch.int32).pin_memory()
            self._ring_i = 0
            self.state = torch.zeros(4, **i32)
            self.pos_arr = torch.zeros(n_slots, **i32)
            # [0] error bits (1 hop watchdog, 2 dependency watchdog, 4 sampler overflow), [1] aborted (poison seen /
            # own watchdog), [2:4] 64-bit exposed-wait cycle counter — see csrc/common.cuh
            self.status = torch.zeros(4, **i32)
            self.done_ctr = torch.zeros(1, **i32)
            # intra-stage flag dependencies (see common.cuh: dep_wait / dep_signal): one flag per kernel of a step
            self.dep_flags = torch.zeros(1024, **i32)
            self.dep_ctr = torch.zeros(17 * 32, **i32)  # two-level ticket counters, one L2 line each
            self._step_seq = 0  # host mirror of ctx[STEP] (device mode: advance_step counts the same steps)
            self.xa = torch.zeros(C, **bf)
            self.xb = torch.zeros(C, **bf)
            self.xc = torch.zeros(C, **bf)  # third residual buffer: parallel-residual blocks keep x, x + attn and the output apart
            self.q = torch.zeros(cfg.n_head * cfg.head_size, **bf)
            self.y_attn = torch.zeros(cfg.n_head * cfg.head_size, **bf)
            self.h_mlp = torch.zeros(cfg.intermediate_size, **bf)
            self.moe = cfg.mlp_class_name == "LLaMAMoE"
            if self.moe:
                # the router's per-token output (expert ids, routing weights), the running sum of the experts' outputs
                # (ping-pong) and, per block, device tables of the experts' weight pointers (csrc/decode_linear.cu)
                self.moe_sel = torch.zeros(8, **i32)
                self.moe_wts = torch.zeros(8, dtype=torch.float32, device=dev)
                self.moe_acc = [torch.zeros(C, **bf), torch.zeros(C, **bf)]
                self.moe_ptrs: Dict[int, Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = {}
                for li, blk in enumerate(model.transformer.h):
                    if getattr(blk, "has_mlp", True) and hasattr(blk, "mlp"):
                        self.moe_ptrs[li] = tuple(  # type: ignore[assignment]
                            torch.tensor([getattr(e, n).weight.data_ptr() for e in blk.mlp.experts], dtype=torch.int64, device=dev)
                            for n in ("fc_1", "fc_2", "proj"))
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            # split-KV spans per KV group: ~2 CTAs per SM over all groups, a multiple of 8 so that the spans of a group
            # launch as clusters of 8 (short contexts merge through distributed shared memory, decode_attention.cu)
            self.n_split = max(1, min(64, (2 * sms) // max(1, cfg.n_query_groups)))
            if self.n_split >= 8:
                self.n_split = (self.n_split + 7) // 8 * 8
            self.part = torch.zeros(cfg.n_head * self.n_split * (cfg.head_size + 2), dtype=torch.float32, device=dev)
            self.tickets = torch.zeros(cfg.n_query_groups, **i32)
            if self.is_starter:
                self.logits = torch.zeros(cfg.padded_vocab_size, dtype=torch.float32, device=dev)
                self.tokens = torch.zeros(n_slots, self.S + 1, **i32)
                self.last_token = torch.zeros(n_slots, **i32)
                self.sample_scratch = ops.sample_scratch(dev)
                # device timeline: %globaltimer of every sampled token + the generation's time base
                self.tok_ts = torch.zeros(n_slots, self.S + 1, dtype=torch.int64, device=dev)
                self.t0_ts = torch.zeros(1, dtype=torch.int64, device=dev)
        self.hop_self = HopTarget(self.hidden_in.data_ptr(), self.flags.data_ptr())
        self._last_x: Optional[Tuple[torch.Tensor, int]] = None
        self._graphs: Dict[Any, ops.CudaGraph] = {}
        self._trace: Optional[torch.Tensor] = None  # device tracer records [n, 6] int64 (see common.cuh)
        self._trace_names: List[str] = []
        self._check_weights()
        self.weight_dtype = weight_dtype
        # prefill attention: "tcgen05" (hand-written flash attention, prompts start at position 0) or "sdpa"
        self.prefill_attn = os.environ.get("MDI_PREFILL_ATTN", "tcgen05")
        self.fused_prefill = fused_prefill_supports(cfg)
        # decode hop: last-CTA row copy (one system-scope fence per step) or per-CTA remote stores + fences
        self.hop_copy = os.environ.get("MDI_HOP_COPY", "1") != "0"
        # MB of the MLP's gate / up weights that the attention output projection pulls into L2 while it waits for
        # the (latency-bound) attention kernel, and chunk pairs per warp of its own rows beyond the ring
        self.pf_next_mb = float(os.environ.get("MDI_PF_NEXT_MB", "0"))
        self.pf_self_chunks = int(os.environ.get("MDI_PF_SELF_CHUNKS", "0"))
        self._q: Dict[int, Tuple[torch.Tensor, torch.Tensor]] = {}  # id(linear) -> (fp8 weight, block scales)
        self._qt: Dict[int, torch.Tensor] = {}  # id(linear) -> transposed block scales [K/128, N] for the prefill GEMM
        if weight_dtype == "fp8":
            self._quantize(free_bf16)
        elif weight_dtype != "bf16":
            raise ValueError("weight_dtype must be 'bf16' or 'fp8'")

    # ---- weights ---------------------------------------------------------------------------------
    def _check_weights(self) -> None:
        for n, p in self.model.named_parameters():
            if p.dtype != torch.bfloat16 or not p.is_contiguous() or p.device != self.device:
                raise ValueError(f"parameter {n}: need contiguous bf16 on {self.device}")

    def _quantize(self, free_bf16: bool) -> None:
        """fp8-e4m3 block-scaled copies of every projection (BASELINE config #5); embeddings/norms stay bf16."""
        from ..utils.quantize import quantize_fp8_block

        if self.moe:
            raise ValueError("fp8 weights are not available for mixture-of-experts models on the fused engine")
        lins = []
        for blk in self.model.transformer.h:
            if getattr(blk, "has_attn", True):
                lins += [blk.attn.attn, blk.attn.proj]
            mlp = getattr(blk, "mlp", None)  # absent in a block cut down to its attention
            if mlp is None:
                continue
            if hasattr(mlp, "fc"):  # plain two-matrix MLP
                lins += [mlp.fc, mlp.proj]
                continue
            if getattr(blk, "has_gu", True):
                lins += [mlp.fc_1, mlp.fc_2]
            if getattr(blk, "has_down", True):
                lins.append(mlp.proj)
        if self.is_starter and not self.cfg.tie_embeddings:
            lins.append(self.model.lm_head)
        with torch.cuda.device(self.device):
            for lin in lins:
                self._q[id(lin)] = quantize_fp8_block(lin.weight.data)
                self._qt[id(lin)] = self._q[id(lin)][1].t().contiguous()
                if free_bf16:
                    lin.weight.data = torch.empty(0, dtype=torch.bfloat16, device=self.device)

    def _w(self, lin: Any, second: bool = False) -> Dict[str, Any]:
        """kwargs selecting a projection's weights for ``ops.linear_decode`` / ``qkv_decode``."""
        q = self._q.get(id(lin))
        if second:
            return dict(W2=lin.weight, bias2=lin.bias) if q is None else dict(W2=q[0], wscale2=q[1], bias2=lin.bias)
        return dict(W=lin.weight, bias=lin.bias) if q is None else dict(W=q[0], wscale=q[1], bias=lin.bias)

    def _gemm(self, a: torch.Tensor, lin: Any, lin2: Any = None, **kw: Any) -> Optional[torch.Tensor]:
        """One prefill projection on the tensor cores: bf16 weights -> ``ops.gemm`` (kind::f16); fp8 block-scaled
        weights -> activations quantised per token and ``ops.gemm_fp8`` (kind::f8f6f4, scales folded in per
        128-element K block) — the fp8 checkpoint is never expanded to bf16.  ``lin2``: gated MLP in one pass."""
        q = self._q.get(id(lin))
        if q is None:
            if lin2 is not None:
                kw.update(w2=lin2.weight, bias2=lin2.bias)
            return ops.gemm(a, lin.weight, bias=lin.bias, **kw)
        kw.pop("block_n", None)
        a8, a_s = ops.quantize_rows_fp8(a)
        if lin2 is not None:
            q2 = self._q[id(lin2)]
            kw.update(w2_8=q2[0], w2_scale_t=self._qt[id(lin2)], bias2=lin2.bias)
        return ops.gemm_fp8(a8, a_s, q[0], self._qt[id(lin)], bias=lin.bias, **kw)

    def _ctas(self, kernel: str) -> int:
        """CTAs per SM for one of the decode linears; ``MDI_CTAS_<KERNEL>`` (e.g. ``MDI_CTAS_GATE_UP=-256``:
        negative = absolute grid size) overrides the stage-wide setting for tuning experiments."""
        v = os.environ.get(f"MDI_CTAS_{kernel.upper()}")
        return int(v) if v else self.ctas_per_sm

    def _variant(self, kernel: str) -> Dict[str, int]:
        """Weight-streaming path of one decode linear: ``MDI_VARIANT_<KERNEL>`` (0 LDG, 1 ring x4 / 1 CTA per SM,
        2 ring x2, 3 ring x3) overrides the library default — small matrices want a deeper ring (the whole row pair
        in flight before the input arrives), the big MLP matrices want more resident warps."""
        v = os.environ.get(f"MDI_VARIANT_{kernel.upper()}")
        return {"variant": int(v)} if v else {}

    def _norm(self, norm: Any) -> Dict[str, Any]:
        """Fused-prologue arguments for a norm module (RMSNorm, Gemma's unit-offset RMSNorm or LayerNorm)."""
        if isinstance(norm, torch.nn.LayerNorm):
            return dict(norm_w=norm.weight, norm_b=norm.bias, layer_norm=True, eps=norm.eps)
        return dict(norm_w=norm.weight, eps=self.cfg.norm_eps, unit_offset=self.cfg.unit_offset_norm)

    def _plain_act(self) -> str:
        return "gelu_tanh" if self.cfg.gelu_approximate == "tanh" else "gelu_erf"

    def _gate_act(self) -> str:
        if self.cfg.mlp_class_name in ("LLaMAMLP", "LLaMAMoE"):  # SwiGLU (the experts of a mixture are LLaMAMLPs)
            return "silu_gate"
        return "gelu_tanh_gate" if self.cfg.gelu_approximate == "tanh" else "gelu_erf_gate"

    # ---- device-side tracing (per-kernel timeline of a stage step) ---------------------
    def _tr(self, name: str) -> Optional[int]:
        if self._trace is None:
            return None
        i = len(self._trace_names)
        if i >= self._trace.shape[0]:
            return None
        self._trace_names.append(name)
        return self._trace[i].data_ptr()

    def trace_step(self, builder: Any, max_records: int = 512, detail: Sequence[str] = (),
                   detail_ctas: int = 1024, graph: bool = True) -> List[Dict[str, Any]]:
        """Run ``builder()`` (a sequence of ``enqueue_*`` calls) once with the kernel tracer on and
        return one row per launch: times in µs relative to the first kernel's entry —
        ``entry`` (first CTA starts), ``ready`` (first CTA past the PDL/hop wait), ``staged`` (last CTA
        has its input in shared memory), ``first_exit`` / ``last_exit`` and the CTA count.  Kernels named
        in ``detail`` additionally get a per-CTA table (``row["per_cta"]``: entry, ready, staged, exit in µs
        and the SM each CTA ran on)."""
        want = set(detail)
        with torch.cuda.device(self.device):
            self._trace = torch.zeros(max_records, 8, dtype=torch.int64, device=self.device)
            self._trace[:, [0, 1, 3]] = torch.iinfo(torch.int64).max
            self._trace_names = []
            table = None
            if want:
                # one per-CTA table per record, wired up BEFORE the launches so that nothing is enqueued
                # between the traced kernels (extra launches would break the PDL chain being measured)
                table = torch.zeros(max_records, detail_ctas, 8, dtype=torch.int64, device=self.device)
                self._trace[:, 6] = table.data_ptr() + torch.arange(max_records, device=self.device) * (detail_ctas * 64)
                self._trace[:, 7] = detail_ctas
            try:
                if graph:  # replayed as ONE graph: launch gaps are the device's, not the Python caller's
                    g = ops.CudaGraph()
                    with g:
                        builder()
                    g.launch()
                else:
                    builder()
                torch.cuda.synchronize(self.device)
                rec = self._trace[: len(self._trace_names)].cpu()
                names = list(self._trace_names)
                details = {}
                if table is not None:
                    for i, n in enumerate(names):
                        if n in want or "*" in want:
                            details[n] = table[i].cpu()
            finally:
                names, self._trace, self._trace_names = list(self._trace_names), None, []
        t0 = int(rec[:, 0].min()) if len(names) else 0
        rows = []
        for i, n in enumerate(names):
            e, r, st_, fx, lx, c = (int(x) for x in rec[i][:6])
            rows.append({"kernel": n, "entry": (e - t0) / 1e3, "ready": (r - t0) / 1e3, "staged": (st_ - t0) / 1e3 if st_ else None,
                         "first_exit": (fx - t0) / 1e3, "last_exit": (lx - t0) / 1e3, "ctas": c})
            if n in details:
                d = details[n][: min(c, details[n].shape[0])]
                rows[-1]["per_cta"] = [{"cta": j, "sm": int(x[5]), "entry": (int(x[0]) - t0) / 1e3, "ready": (int(x[1]) - t0) / 1e3,
                                        "staged": (int(x[2]) - t0) / 1e3, "exit": (int(x[4]) - t0) / 1e3,
                                        "p6": (int(x[6]) - t0) / 1e3 if int(x[6]) else None} for j, x in enumerate(d)]
        return rows

    # ---- kernel sequences ------------------------------------------------------------------------
    def enqueue_head(self, wait: bool, stats: bool = True) -> None:
        """starter: final RMSNorm + lm_head on ``hidden_in[slot]`` → fp32 logits (+ the sampler's
        logit histogram / arg-max, gathered in the same pass when ``stats``)."""
        m, cfg = self.model, self.cfg
        lw = self._w(m.lm_head)
        ops.linear_decode(
            lw.pop("W"), self.hidden_in, self.logits, self.ctx, **lw, **self._norm(m.transformer.ln_f),
            x_slot_stride=cfg.n_embd, wait_flag=self.flags.data_ptr() if wait else None,
            status=self.status.data_ptr(), wait_max_cycles=self.wait_max_cycles, ctas_per_sm=self._ctas("lm_head"),
            use_pdl=self.use_pdl, stats=self.sample_scratch if stats else None, trace=self._tr("lm_head"))

    def enqueue_sample(self) -> None:
        """Must follow ``enqueue_head(stats=True)``: consumes (and clears) the logit statistics."""
        s = self.sampling
        greedy = not (s.temperature > 0.0 or s.top_p > 0.0)
        ops.sample_fast(self.logits, self.sample_scratch, self.tokens, self.ctx, vocab=self.cfg.padded_vocab_size,
                        top_k=s.top_k, temperature=s.temperature, greedy=greedy,
                        seed=s.seed if s.seed is not None else 0x5EED, tok_slot_stride=self.tokens.shape[1],
                        last_token=self.last_token, use_pdl=self.use_pdl, top_p=1.0 if greedy else s.top_p,
                        tok_ts=self.tok_ts, status=self.status)

    def enqueue_embed(self, from_tokens: bool) -> None:
        m, cfg = self.model, self.cfg
        ops.embed(m.transformer.wte.weight, self.xa, self.ctx, tokens=self.tokens if from_tokens else None,
                  tok_slot_stride=self.tokens.shape[1], scale=float(cfg.n_embd ** 0.5) if cfg.scale_embeddings else 1.0,
                  wpe=m.transformer.wpe.weight if "wpe" in m.transformer else None, use_pdl=self.use_pdl)

    def _units(self) -> List[Tuple[int, str]]:
        """The stage's work as residual sub-layers ``(local block, "attn" | "mlp" | "gu" | "down")`` — whole
        blocks contribute attn + mlp; a sub-layer pipeline boundary leaves a partial block (models/partition.py)."""
        out: List[Tuple[int, str]] = []
        for li, blk in enumerate(self.model.transformer.h):
            if self.cfg.parallel_residual:
                out.append((li, "par"))  # attention and MLP both read the block input: one indivisible unit
                continue
            if getattr(blk, "has_attn", True):
                out.append((li, "attn"))
            gu, down = getattr(blk, "has_gu", True), getattr(blk, "has_down", True)
            if gu and down:
                out.append((li, "mlp"))
            elif gu:
                out.append((li, "gu"))    # stage ends after the gate/up projections: the hop carries [x | h]
            elif down:
                out.append((li, "down"))  # stage starts at the down projection of a layer cut after gate/up
        return out

    def enqueue_blocks(self, hop: Optional[HopTarget], wait_input: bool, dep_flags: bool = False) -> None:
        """All local sub-layers for one token.  Input residual: ``xa`` on the starter (embedding),
        ``hidden_in[slot]`` on a secondary.  The first kernel (QKV projection; gate/up or down projection when
        the stage starts inside a layer) acquires the incoming hop; the last one (down projection; attention
        output projection or gate/up when the stage ends inside a layer) finishes the outgoing row in
        ``out_local[slot]`` and — with a ``hop`` target — its last CTA copies the row into the next stage's
        ``hidden_in[slot]`` over NVLink and releases the flag.  A stage that ends after gate/up sends
        ``[x | h]``; the stage that starts at the matching down projection reads both halves from its
        ``hidden_in`` row."""
        cfg, C = self.cfg, self.cfg.n_embd
        I = cfg.intermediate_size
        W_in, W_out = self.W_in, self.W_out
        x_in, x_in_stride = (self.xa, 0) if self.is_starter else (self.hidden_in, W_in)
        common = dict(use_pdl=self.use_pdl)
        units = self._units()
        n_moe = 1 + 2 * cfg.n_expert_per_token if self.moe else 2  # router + (gate/up, down) per chosen expert
        n_kernels = sum({"attn": 3, "mlp": n_moe, "gu": 1, "down": 1, "par": 5}[k] for _, k in units)
        if self.moe and dep_flags:
            raise RuntimeError("flag dependencies are not wired through the mixture-of-experts kernels (use PDL)")
        kidx = [0]
        st_kw = dict(status=self.status.data_ptr(), wait_max_cycles=self.wait_max_cycles)

        def dep() -> Dict[str, Any]:
            """Flag dependency wiring of the next launch: wait on the previous kernel's flag, publish ours
            (the first kernel keeps the grid-level wait, the last one has no local consumer)."""
            j = kidx[0]
            kidx[0] += 1
            if not dep_flags:
                return {}
            base = self.dep_flags.data_ptr()
            return dict(dep_wait=base + 4 * (j - 1) if j > 0 else None,
                        dep_signal=base + 4 * j if j < n_kernels - 1 else None, dep_ctr=self.dep_ctr.data_ptr())

        def out_kw(name: str, n_pre: int = 0, pre: Optional[Tuple[torch.Tensor, int]] = None) -> Dict[str, Any]:
            """Destination of the stage's LAST kernel: the local row (offset ``n_pre`` elements when ``[x | h]`` is
            sent) and, with a hop target, the copy + signal done by its last CTA."""
            kw: Dict[str, Any] = dict(y_ptr=self.out_local.data_ptr() + 2 * n_pre, y_slot_stride=W_out, trace=self._tr(name))
            if hop is None:
                return kw
            if not self.hop_copy:
                if n_pre:
                    raise RuntimeError("stages that end after a gate/up unit need the row-copy hop (MDI_HOP_COPY=1)")
                return dict(y_ptr=hop.hidden_ptr, y_slot_stride=W_out, signal_flag=hop.flag_ptr,
                            done_ctr=self.done_ctr.data_ptr(), trace=self._tr(name + "+hop"))
            kw.update(hop_ptr=hop.hidden_ptr, hop_slot_stride=W_out, signal_flag=hop.flag_ptr, done_ctr=self.done_ctr.data_ptr(),
                      trace=self._tr(name + "+hop"))
            if pre is not None:
                kw.update(hop_pre=(pre[0].data_ptr(), pre[1], n_pre))
            return kw

        self._last_x = None
        plain_mlp = cfg.mlp_class_name == "GptNeoxMLP"

        def attn_part(li: int, blk: Any, x_src: torch.Tensor, x_stride: int, first: bool, wait: Dict[str, Any]) -> None:
            """QKV projection (+norm, RoPE, KV append) and split-KV attention of block ``li`` -> ``self.y_attn``."""
            kv_layer = self.kv[li]
            qw = self._w(blk.attn.attn)
            ops.qkv_decode(
                qw.pop("W"), x_src, self.model.cos, self.model.sin, self.q, kv_layer, self.ctx,
                n_head=cfg.n_head, n_groups=cfg.n_query_groups, head_size=cfg.head_size,
                rope_n_elem=cfg.rope_n_elem, max_seq=self.S, **self._norm(blk.norm_1), **qw, x_slot_stride=x_stride,
                trace=self._tr(f"L{li}.qkv"), ctas_per_sm=self._ctas("qkv"), ctx_early=not first, **self._variant("qkv"),
                **wait, **dep(), **common)
            ops.attn_decode(self.q, kv_layer, self.y_attn, self.part, self.tickets, self.ctx, n_head=cfg.n_head,
                            n_groups=cfg.n_query_groups, head_size=cfg.head_size, max_seq=self.S,
                            n_split=self.n_split, use_pdl=self.use_pdl, trace=self._tr(f"L{li}.attn"), **st_kw, **dep())

        def mlp_up(li: int, blk: Any, x_src: torch.Tensor, x_stride: int, norm: Any, first: bool, wait: Dict[str, Any],
                   extra: Optional[Dict[str, Any]] = None) -> None:
            """First MLP pass -> ``self.h_mlp`` (or the destination in ``extra``): gate/up with act(g) * u, or fc + GELU."""
            kw = dict(**self._norm(norm), x_slot_stride=x_stride, ctas_per_sm=self._ctas("gate_up"), ctx_early=not first,
                      **(wait if first else st_kw), **dep(), **common)
            dst = extra if extra is not None else dict(trace=self._tr(f"L{li}.{'fc' if plain_mlp else 'gate_up'}"))
            y = None if extra is not None else self.h_mlp
            if plain_mlp:
                fw = self._w(blk.mlp.fc)
                ops.linear_decode(fw.pop("W"), x_src, y, self.ctx, **fw, act=self._plain_act(), **kw, **dst)
            else:
                gw = {**self._w(blk.mlp.fc_1), **self._w(blk.mlp.fc_2, second=True)}
                ops.linear_decode(gw.pop("W"), x_src, y, self.ctx, **gw, act=self._gate_act(), **kw, **dst)

        def moe_mlp(li: int, blk: Any, x_src: torch.Tensor, x_stride: int, wait: Dict[str, Any], dst: Optional[torch.Tensor],
                    last: bool) -> None:
            """Routed MLP (model.py:823-853) without a host round trip: the router kernel leaves the token's expert ids
            and routing weights on the device; each chosen expert is a gate/up + a down launch that looks its weight
            pointers up after the dependency wait.  The last down pass adds the residual and, on the stage's last
            block, finishes the outgoing row (local row + hop copy)."""
            top = cfg.n_expert_per_token
            p1, p2, p3 = self.moe_ptrs[li]
            norm = self._norm(blk.norm_2)
            ops.moe_router(blk.mlp.gate.weight, x_src, self.moe_sel, self.moe_wts, self.ctx, top=top, **norm,
                           x_slot_stride=x_stride, trace=self._tr(f"L{li}.router"), **wait, **common)
            kidx[0] += 1
            for k in range(top):
                ops.moe_linear_decode(p1, x_src, self.h_mlp, self.ctx, self.moe_sel, self.moe_wts, k, N=I, K=C, w2_ptrs=p2,
                                      **norm, act=self._gate_act(), x_slot_stride=x_stride, ctas_per_sm=min(3, self._ctas("gate_up")),
                                      status=self.status.data_ptr(), trace=self._tr(f"L{li}.e{k}.gate_up"), sel_early=k > 0, **common)
                # every pass but the first runs >= 2 launches after the router: its expert is known before the PDL wait
                kw: Dict[str, Any] = dict(N=C, K=I, prev=self.moe_acc[(k + 1) % 2] if k > 0 else None, sel_early=True,
                                          ctas_per_sm=min(3, self._ctas("down")), status=self.status.data_ptr(), **common)
                if k < top - 1:
                    ops.moe_linear_decode(p3, self.h_mlp, self.moe_acc[k % 2], self.ctx, self.moe_sel, self.moe_wts, k,
                                          trace=self._tr(f"L{li}.e{k}.down"), **kw)
                elif not last:
                    ops.moe_linear_decode(p3, self.h_mlp, dst, self.ctx, self.moe_sel, self.moe_wts, k, residual=x_src,
                                          res_slot_stride=x_stride, trace=self._tr(f"L{li}.e{k}.down"), **kw)
                else:
                    ops.moe_linear_decode(p3, self.h_mlp, None, self.ctx, self.moe_sel, self.moe_wts, k, residual=x_src,
                                          res_slot_stride=x_stride, **out_kw(f"L{li}.e{k}.down"), **kw)
                kidx[0] += 2

        def free_buf(*busy: Any) -> torch.Tensor:
            return next(b_ for b_ in (self.xa, self.xb, self.xc) if all(b_ is not o for o in busy))

        for ui, (li, kind) in enumerate(units):
            first, last = ui == 0, ui == len(units) - 1
            blk = self.model.transformer.h[li]
            wait = dict(wait_flag=self.flags.data_ptr() if (first and wait_input and not self.is_starter) else None, **st_kw)
            x_out = free_buf(x_in)
            pf: Dict[str, Any] = {}
            if kind == "par":
                # parallel residual (model.py:596-629): out = x + attn(norm_1 x) + mlp(norm_2 x | norm_1 x)
                attn_part(li, blk, x_in, x_in_stride, first, wait)
                ow = self._w(blk.attn.proj)
                x_mid = x_out
                ops.linear_decode(ow.pop("W"), self.y_attn, x_mid, self.ctx, **ow, residual=x_in, res_slot_stride=x_in_stride,
                                  ctx_early=True, ctas_per_sm=self._ctas("o_proj"), trace=self._tr(f"L{li}.o_proj"), **st_kw,
                                  **dep(), **common)
                mlp_up(li, blk, x_in, x_in_stride, blk.norm_1 if blk.norm_2 is None else blk.norm_2, False, wait)
                x_out = free_buf(x_in, x_mid)
                lw, src, name, ctas = self._w(blk.mlp.proj), dict(x=self.h_mlp), f"L{li}.down", self._ctas("down")
                res = dict(residual=x_mid, res_slot_stride=0, ctx_early=True, ctas_per_sm=ctas, **st_kw, **dep())
                w_out = lw.pop("W")
                if not last:
                    ops.linear_decode(w_out, src.pop("x"), x_out, self.ctx, **lw, **res, trace=self._tr(name), **common)
                    x_in, x_in_stride = x_out, 0
                else:
                    ops.linear_decode(w_out, src.pop("x"), None, self.ctx, **lw, **res, **out_kw(name), **common)
                continue
            if kind == "attn":
                attn_part(li, blk, x_in, x_in_stride, first, wait)
                lw, src, name, ctas = self._w(blk.attn.proj), dict(x=self.y_attn), f"L{li}.o_proj", self._ctas("o_proj")
                pf.update(self._variant("o_proj"))
                if self.pf_self_chunks:
                    pf["l2_pf_chunks"] = self.pf_self_chunks
                if self.pf_next_mb > 0 and getattr(blk, "has_gu", False) and not plain_mlp and not self.moe:
                    w1, w2 = self._w(blk.mlp.fc_1)["W"], self._w(blk.mlp.fc_2)["W"]
                    nbytes = min(int(self.pf_next_mb * 2 ** 20) // 2, w1.numel() * w1.element_size()) & ~4095
                    pf["prefetch"] = (w1.data_ptr(), w2.data_ptr(), nbytes)
            if kind == "gu":  # the stage ends after the gate/up pass: out row = [x | h], x copied by the hop's last CTA
                assert last, "a gate/up-only unit is the last unit of its stage"
                self._last_x = (x_in, x_in_stride)
                mlp_up(li, blk, x_in, x_in_stride, blk.norm_2, first, wait,
                       extra=out_kw(f"L{li}.gate_up", n_pre=C, pre=(x_in, x_in_stride)))
                return
            if kind == "mlp" and self.moe:
                moe_mlp(li, blk, x_in, x_in_stride, wait, None if last else x_out, last)
                x_in, x_in_stride = x_out, 0
                continue
            if kind == "mlp":
                mlp_up(li, blk, x_in, x_in_stride, blk.norm_2, first, wait)
                lw, src, name, ctas = self._w(blk.mlp.proj), dict(x=self.h_mlp), f"L{li}.down", self._ctas("down")
                pf.update(self._variant("down"))
            if kind == "down":  # the stage starts here: h and the residual are the two halves of the incoming row
                assert first and not self.is_starter, "a down-only unit is the first unit of a secondary stage"
                lw, name, ctas = self._w(blk.mlp.proj), f"L{li}.down", self._ctas("down")
                src = dict(x=None, x_ptr=self.hidden_in.data_ptr() + 2 * C, x_slot_stride=W_in)
                res = dict(residual=None, residual_ptr=self.hidden_in.data_ptr(), res_slot_stride=W_in, ctx_early=False,
                           ctas_per_sm=ctas, **wait, **dep())
            else:
                res = dict(residual=x_in, res_slot_stride=x_in_stride, ctx_early=True,  # never the first launch after advance_step
                           ctas_per_sm=ctas, **st_kw, **pf, **dep())
            w_out = lw.pop("W")
            if not last:
                ops.linear_decode(w_out, src.pop("x"), x_out, self.ctx, **src, **lw, **res, trace=self._tr(name), **common)
                x_in, x_in_stride = x_out, 0
            else:
                ops.linear_decode(w_out, src.pop("x"), None, self.ctx, **src, **lw, **res, **out_kw(name), **common)

    # ---- prefill (T > 1): linears on the tcgen05 GEMM ------------------------------------------------
    @torch.inference_mode()
    def prefill(self, data: torch.Tensor, input_pos: torch.Tensor, slot: int,
                hop: Optional[Tuple[int, int]] = None) -> Optional[torch.Tensor]:
        """All local blocks for a whole prompt.  Every projection (SURVEY K3/K8/K10/K11 at T > 1) runs
        on the hand-written tcgen05/TMEM/TMA GEMM with its bias/residual epilogue (gate and up
        projections in one dual-accumulator GEMM with the SiLU/GELU·mul epilogue); RMSNorm is the row
        kernel; RoPE, the KV-slot write and causal attention reuse the eager helper.
        ``data``: token ids ``[1,T]`` on the starter, hidden state ``[1,T,C]`` on a secondary.
        ``hop = (dst_ptr, flag_ptr)``: the last down-projection's epilogue stores its ``[T,C]`` output
        (residual added) straight into the next stage's buffer at ``dst_ptr`` (peer memory) and releases
        ``flag[ctx.slot] = ctx.signal`` from inside the GEMM — the fused prefill hop; returns None."""
        m, cfg = self.model, self.cfg
        C = cfg.n_embd
        T = data.size(1)
        if T > self.S:
            raise ValueError(f"prompt of {T} tokens does not fit the stage's KV slots ({self.S} positions)")
        g_in = None
        if not self.fused_prefill:
            # architectures outside the tcgen05 prefill subset (LayerNorm / parallel residual / plain MLP / learned
            # positions): the prompt goes through the eager modules on the SAME KV pool, decode stays fused
            out = m(data.long() if self.is_starter else data.to(torch.bfloat16), input_pos, slot=slot).to(torch.bfloat16).contiguous()
            if hop is None:
                return out
            ops.check(ops.lib().mdi_copy_signal(out.data_ptr(), hop[0], out.numel() * 2, hop[1], self.done_ctr.data_ptr(),
                                                self.ctx.data_ptr(), self.status.data_ptr(), ops.stream_ptr()), "prefill hop (eager)")
            self._keep = out  # the copy kernel reads it asynchronously
            return None
        if self.is_starter:
            x = m.embed(data.long(), input_pos)[0].to(torch.bfloat16).contiguous()
        elif self.W_in > C:  # [x | h]: the stage starts at the down projection of a layer cut after gate/up
            d = data[0].to(torch.bfloat16)
            x, g_in = d[:, :C].contiguous(), d[:, C:].contiguous()
        else:
            x = data[0].to(torch.bfloat16).contiguous()
        cos, sin = m.rope_for(T, input_pos)
        eps, uo = cfg.norm_eps, cfg.unit_offset_norm
        units = self._units()
        # tile choice: prompts of >= 192 rows take the CTA-pair kernel (cta_group::2: two CTAs share a 256 x 256 MMA tile,
        # 86-97 % of cuBLAS on the Llama-3 shapes and ahead of it on the fused gated MLP); shorter ones want more,
        # smaller tiles
        pair = T >= 192 and os.environ.get("MDI_GEMM_PAIR", "1") != "0"
        bn = 512 if pair else (256 if T > 128 else 128)
        # the tcgen05 attention kernel covers a prompt that starts at position 0 (every prefill of the pipeline);
        # anything else (a continuation at an offset) takes the eager helper
        use_fa = (self.prefill_attn == "tcgen05" and cfg.rope_n_elem % 16 == 0
                  and int(input_pos[0]) == 0 and int(input_pos.numel()) == T)

        def out_gemm(a_in: torch.Tensor, lin: Any, last: bool) -> Optional[torch.Tensor]:
            if hop is not None and last:
                self._gemm(a_in, lin, residual=x, out_ptr=hop[0], signal_flag=hop[1], done_ctr=self.done_ctr, ctx=self.ctx,
                           status=self.status, block_n=bn)
                return None
            return self._gemm(a_in, lin, residual=x, block_n=bn)

        for ui, (li, kind) in enumerate(units):
            blk, last = m.transformer.h[li], ui == len(units) - 1
            if kind == "attn":
                h = ops.rmsnorm_rows(x, blk.norm_1.weight, eps, uo)
                qkv = self._gemm(h, blk.attn.attn, block_n=bn)
                if use_fa:  # RoPE + KV append + causal flash attention with S / P.V in TMEM
                    y = ops.attn_prefill(qkv, m.cos, m.sin, self.kv[li], slot, n_head=cfg.n_head,
                                         n_groups=cfg.n_query_groups, head_size=cfg.head_size, rope_n_elem=cfg.rope_n_elem)
                else:  # eager helper (SDPA): the oracle path
                    y = blk.attn.attend_qkv(qkv.unsqueeze(0), cos, sin, input_pos, m.kv_pool.layer(li, slot))[0].contiguous()
                x = out_gemm(y, blk.attn.proj, last)
            elif kind == "down":
                x = out_gemm(g_in, blk.mlp.proj, last)
            elif self.moe:
                x = self._moe_prefill(blk, x, eps, uo)
                if last and hop is not None:
                    ops.check(ops.lib().mdi_copy_signal(x.data_ptr(), hop[0], x.numel() * 2, hop[1], self.done_ctr.data_ptr(),
                                                        self.ctx.data_ptr(), self.status.data_ptr(), ops.stream_ptr()),
                              "prefill hop (mixture of experts)")
                    self._keep = x  # the copy kernel reads it asynchronously
                    return None
            else:
                h = ops.rmsnorm_rows(x, blk.norm_2.weight, eps, uo)
                g = self._gemm(h, blk.mlp.fc_1, blk.mlp.fc_2, act=self._gate_act(), **({"block_n": 512} if pair else {}))
                if kind == "gu":  # the stage ends here: [x | h] travels (copy + in-kernel flag release)
                    out = torch.cat((x, g), dim=1)
                    if hop is None:
                        return out.unsqueeze(0)
                    ops.check(ops.lib().mdi_copy_signal(out.data_ptr(), hop[0], out.numel() * 2, hop[1], self.done_ctr.data_ptr(),
                                                        self.ctx.data_ptr(), self.status.data_ptr(), ops.stream_ptr()),
                              "prefill hop ([x | h])")
                    return None
                x = out_gemm(g, blk.mlp.proj, last)
            if x is None:
                return None
        return x.unsqueeze(0)

    def _moe_prefill(self, blk: Any, x: torch.Tensor, eps: float, uo: bool) -> torch.Tensor:
        """Routed MLP of a whole prompt (model.py:823-853) on the tcgen05 GEMMs: route every token (top-k of the bf16
        router logits, softmax over the chosen ones), sort the (token, expert) pairs by expert, and run each expert's
        gated gate/up GEMM and down GEMM over ITS rows only.  One host read per layer (the per-expert row counts shape
        the GEMM launches; prompt processing is not graph-captured) against the reference's ``torch.where`` per expert."""
        mlp, top = blk.mlp, self.cfg.n_expert_per_token
        T = x.shape[0]
        h = ops.rmsnorm_rows(x, blk.norm_2.weight, eps, uo)
        weight, chosen = torch.topk(mlp.gate(h), top, dim=-1)  # [T, top]; the router is E x C: negligible
        weight = weight.softmax(dim=-1, dtype=torch.float).to(x.dtype)
        flat = chosen.reshape(-1)
        order = torch.argsort(flat, stable=True)
        counts = torch.bincount(flat, minlength=self.cfg.n_expert).tolist()
        tok = order // top
        rows = h.index_select(0, tok)  # [T * top, C], grouped by expert
        w_sorted = weight.reshape(-1).index_select(0, order)
        contrib = torch.empty(T * top, self.cfg.n_embd, dtype=x.dtype, device=x.device)
        off = 0
        for e, n in enumerate(counts):
            if n == 0:
                continue
            ex = mlp.experts[e]
            bn = 512 if n >= 192 and os.environ.get("MDI_GEMM_PAIR", "1") != "0" else (256 if n > 128 else 128)
            g = self._gemm(rows[off:off + n], ex.fc_1, ex.fc_2, act="silu_gate", block_n=bn)
            self._gemm(g, ex.proj, out=contrib[off:off + n], block_n=bn)
            off += n
        contrib.mul_(w_sorted.unsqueeze(1))
        out = torch.zeros_like(x)
        out.index_add_(0, tok, contrib)  # bf16 accumulation like the eager module
        return x + out

    # ---- graphs ------------------------------------------------------------------------------------
    def graph(self, key: Any, builder: Any, warm: bool = True) -> ops.CudaGraph:
        """Capture ``builder()`` once per ``key``.  ``warm`` runs it eagerly first (loads the
        kernels, sets smem attributes); graphs with hop waits/signals must pass ``warm=False``
        (their side effects are not idempotent) after :meth:`warmup`."""
        g = self._graphs.get(key)
        if g is None:
            with torch.cuda.device(self.device):
                if warm:
                    builder()
                torch.cuda.current_stream().synchronize()
                g = ops.CudaGraph()
                with g:
                    builder()
            self._graphs[key] = g
        return g

    def wait_cycles(self) -> int:
        """Cycles CTA 0 of the hop-consuming kernels spent spinning on the incoming flag so far."""
        lo, hi = self.status[2].item(), self.status[3].item()
        return (lo & 0xFFFFFFFF) | ((hi & 0xFFFFFFFF) << 32)

    def warmup(self) -> None:
        """Launch every kernel of the step once with hops disabled (slot 0, position 0)."""
        with torch.cuda.device(self.device):
            self.set_ctx(0, 0)
            if self.is_starter:
                self.enqueue_head(wait=False)
                self.enqueue_sample()
                self.enqueue_embed(from_tokens=True)
            self.enqueue_blocks(None, False)
            if self.is_starter:
                self.tokens.zero_()
                self.last_token.zero_()
            self.kv[:, 0, :, :, 0].zero_()
            torch.cuda.current_stream().synchronize()

    def set_ctx(self, slot: int, pos: int, wait: int = 0, signal: int = 0, token: int = 0) -> None:
        """Host-driven step descriptor: 32 bytes, pinned host ring entry → device (async).  The
        ring keeps an entry untouched until 4096 later steps were issued, far beyond what can be
        in flight, so the async copy never races with the host."""
        h = self.ctx_ring[self._ring_i]
        self._ring_i = (self._ring_i + 1) % self.ctx_ring.shape[0]
        h[ops.CTX_SLOT], h[ops.CTX_POS], h[ops.CTX_WAIT], h[ops.CTX_SIGNAL], h[ops.CTX_TOKEN] = slot, pos, wait, signal, token
        h[ops.CTX_STEP] = self._step_seq  # same numbering as advance_step's: the flag dependencies wait for step + 1
        self._step_seq += 1
        self.ctx.copy_(h, non_blocking=True)

    def reset_deps(self) -> None:
        """New generation: the step numbering restarts at 0, so the dependency flags do too."""
        self.dep_flags.zero_()
        self.dep_ctr.zero_()
        self._step_seq = 0


class FusedStageRunner(StageRunner):
    """Host-driven adapter: prefill on the eager module, decode through the fused kernels."""

    def __init__(self, model: StageModule, max_seq_length: Optional[int] = None, n_slots: int = 8,
                 sampling: Optional[SamplingParams] = None, use_graphs: bool = True) -> None:
        self.stage = FusedStage(model, n_slots=n_slots, max_seq_length=max_seq_length, sampling=sampling)
        self.model = model.eval()
        self.role = model.role
        self.device = self.stage.device
        self.dtype = torch.bfloat16
        self.slots: Dict[int, int] = {}
        self.use_graphs = use_graphs
        self.n_launches = 0

    def begin_sample(self, sample_id: int) -> None:
        if sample_id in self.slots:
            return
        if len(self.slots) >= self.stage.n_slots:
            raise RuntimeError(f"stage was sized for {self.stage.n_slots} concurrent samples")
        self.slots[sample_id] = len(self.slots)

    def _run(self, key: str, builder: Any) -> None:
        if self.use_graphs:
            self.stage.graph(key, builder).launch()
            self.n_launches += self.stage._graphs[key].n_nodes
        else:
            builder()

    @torch.inference_mode()
    def forward(self, sample_id: int, data: torch.Tensor, input_pos: torch.Tensor) -> torch.Tensor:
        st, slot = self.stage, self.slots[sample_id]
        T = data.size(1)
        with torch.cuda.device(self.device):
            if T > 1:  # prefill: tcgen05 GEMMs + eager attention on the shared KV pool
                return st.prefill(data, input_pos, slot)
            pos = int(input_pos[-1])
            if self.role == "starter":
                st.set_ctx(slot, pos, token=int(data.reshape(-1)[-1]))
                self._run("fwd", lambda: (st.enqueue_embed(from_tokens=False), st.enqueue_blocks(None, False)))
            else:
                st.hidden_in[slot].copy_(data.reshape(-1).to(self.dtype))
                st.set_ctx(slot, pos)
                self._run("fwd", lambda: st.enqueue_blocks(None, False))
            if st._last_x is not None:  # stage ends after gate/up: the outgoing row is [x | h], x added here (no hop kernel)
                buf, stride = st._last_x
                st.out_local[slot, : st.cfg.n_embd].copy_((buf[slot] if stride else buf)[: st.cfg.n_embd])
            return st.out_local[slot].view(1, 1, -1).clone()

    @torch.inference_mode()
    def head(self, hidden: torch.Tensor) -> torch.Tensor:
        st = self.stage
        with torch.cuda.device(self.device):
            st.hidden_in[0].copy_(hidden[0, -1].to(self.dtype))
            st.set_ctx(0, 0)
            self._run("head", lambda: st.enqueue_head(wait=False, stats=False))
            return st.logits.view(1, 1, -1).clone()
