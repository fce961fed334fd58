"""Python bindings of the sm_100a kernel library (``_mdi_ops.so``, C ABI via ctypes).

The wrappers take torch tensors, pass raw device pointers and launch on the *current torch
stream*, so they compose with torch ops and can be captured into CUDA graphs.  There is no
silent fallback: on a CUDA machine a missing/unbuildable library raises, on a CPU-only machine
``available()`` is False and callers use the eager path explicitly.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_size_t, c_ulonglong, c_void_p, POINTER, byref
from pathlib import Path
from typing import Any, Optional, Tuple

import torch

from . import build as _build

__all__ = ["lib", "available", "require", "OpsError", "ptr", "stream_ptr", "check", "ACT", "CTX_INTS",
           "linear_decode", "qkv_decode", "attn_decode", "gemm", "gemm_fp8", "quantize_rows_fp8", "sample_fast", "sample_scratch", "stamp", "POISON", "embed", "rmsnorm_rows", "sample", "advance_step",
           "CudaGraph"]

CTX_SLOT, CTX_POS, CTX_WAIT, CTX_SIGNAL, CTX_TOKEN, CTX_STEP = 0, 1, 2, 3, 4, 5
CTX_INTS = 8
POISON = 0x7FFFFFFF  # flag value of an aborted ring (csrc/common.cuh: MDI_POISON)
ACT = {"none": 0, "silu_gate": 1, "gelu_tanh_gate": 2, "gelu_erf_gate": 3, "gelu_tanh": 4, "gelu_erf": 5}

_lib: Optional[ctypes.CDLL] = None
_load_error: Optional[BaseException] = None


class OpsError(RuntimeError):
    pass


def _declare(lib: ctypes.CDLL) -> None:
    vp, i32, i64, f32 = c_void_p, c_int, c_longlong, c_float
    lib.mdi_error_string.restype = c_char_p
    lib.mdi_error_string.argtypes = [i32]
    lib.mdi_linear_decode.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, i32, i32, f32, i32, i32, i32,
                                      vp, vp, i64, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp,
                                      vp, i64, vp, vp, c_ulonglong, i32, vp, i64, i32, vp, i32, vp]
    lib.mdi_qkv_decode.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, i32, i32, f32, i32,
                                   vp, vp, i64, i32, i32, i32, vp, vp, vp, vp, vp, vp, i32, vp]
    lib.mdi_moe_router.argtypes = [vp, vp, vp, vp, i32, vp, i64, i32, i32, i32, f32, i32, vp, vp, vp, vp, i64, i32, vp, vp]
    lib.mdi_moe_linear_decode.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, vp, i32, vp, vp, vp, i64, i64, i64, i32, i32, f32,
                                          i32, i32, vp, vp, vp, vp, i64, i32, i32, vp, i32, vp]
    lib.mdi_set_moe_variant.argtypes = [i32]
    lib.mdi_set_linear_variant.argtypes = [i32]
    lib.mdi_set_attn_cluster.argtypes = [i32]
    lib.mdi_set_l2_prefetch_mb.argtypes = [i32]
    lib.mdi_get_linear_variant.restype = i32
    lib.mdi_attn_decode.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, i64, vp]
    lib.mdi_embed.argtypes = [vp, vp, vp, i64, vp, vp, i64, i32, f32, i32, vp]
    lib.mdi_rmsnorm_rows.argtypes = [vp, vp, vp, i32, i32, f32, i32, vp]
    lib.mdi_sample.argtypes = [vp, i64, vp, i64, vp, vp, i32, i32, f32, i32, c_ulonglong, i32, f32, vp]
    lib.mdi_sample_fast.argtypes = [vp, vp, vp, i64, vp, vp, i32, i32, f32, i32, c_ulonglong, i32, f32, vp, i64, vp, vp]
    lib.mdi_stamp.argtypes = [vp, vp]
    lib.mdi_sample_scratch_bytes.restype = c_size_t
    lib.mdi_gemm_bf16.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp]
    lib.mdi_gemm_bf16_ex.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp,
                                     i32, i32, i32, i32, vp]
    lib.mdi_gemm_fp8.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp]
    lib.mdi_quantize_rows_fp8.argtypes = [vp, vp, vp, i32, i32, i64, vp]
    lib.mdi_attn_prefill.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp]
    lib.mdi_set_prefill_attn_pipe.argtypes = [i32]
    lib.mdi_get_prefill_attn_pipe.restype = i32
    lib.mdi_advance_step.argtypes = [vp, vp, vp, i32, i32, i32, vp]
    lib.mdi_wait_flag.argtypes = [vp, vp, vp, i64, vp]
    lib.mdi_set_flag.argtypes = [vp, vp, vp]
    lib.mdi_copy_bytes.argtypes = [vp, vp, c_size_t, vp]
    lib.mdi_copy_signal.argtypes = [vp, vp, c_size_t, vp, vp, vp, vp, vp]
    lib.mdi_device_info.argtypes = [POINTER(i32), POINTER(i32), POINTER(i32), POINTER(c_size_t)]
    lib.mdi_p2p_alloc.argtypes = [c_size_t, POINTER(vp), c_char_p]
    lib.mdi_p2p_open.argtypes = [c_char_p, POINTER(vp)]
    lib.mdi_p2p_close.argtypes = [vp]
    lib.mdi_p2p_free.argtypes = [vp]
    lib.mdi_enable_peer.argtypes = [i32, i32]
    lib.mdi_graph_begin.argtypes = [vp]
    lib.mdi_graph_end.argtypes = [vp, POINTER(vp), POINTER(i32)]
    lib.mdi_graph_launch.argtypes = [vp, vp, i32]
    lib.mdi_graph_launch_pattern.argtypes = [vp, i32, vp, i32, i32, vp]
    lib.mdi_graph_destroy.argtypes = [vp]
    lib.mdi_host_alloc.argtypes = [c_size_t, POINTER(vp)]
    lib.mdi_host_free.argtypes = [vp]
    lib.mdi_memcpy_async.argtypes = [vp, vp, c_size_t, i32, vp]
    lib.mdi_stream_sync.argtypes = [vp]
    for name in dir(lib):
        pass


def lib() -> ctypes.CDLL:
    """Load (building if stale and nvcc is present) the kernel library."""
    global _lib, _load_error
    if _lib is not None:
        return _lib
    if _load_error is not None:
        raise OpsError(f"kernel library unavailable: {_load_error}") from _load_error
    try:
        alt = os.environ.get("MDI_OPS_LIB")  # kernel A/B experiments: load another build of the same ABI
        if alt:
            handle = ctypes.CDLL(alt)
        else:
            if not _build.is_fresh():
                if _build.nvcc_path():
                    _build.build()
                elif not _build.LIB.exists():
                    raise OpsError("kernel library not built and nvcc not available")
                else:  # argtypes below describe the CURRENT sources: an older binary would be called with a wrong ABI
                    raise OpsError(f"{_build.LIB} was built from different sources (digest {_build.embedded_digest()}) "
                                   "and nvcc is not available to rebuild it")
            handle = ctypes.CDLL(str(_build.LIB))
        _declare(handle)
        _lib = handle
        if os.environ.get("MDI_PREFILL_ATTN_PIPE"):
            handle.mdi_set_prefill_attn_pipe(int(os.environ["MDI_PREFILL_ATTN_PIPE"]))
        if os.environ.get("MDI_ATTN_CLUSTER"):
            handle.mdi_set_attn_cluster(int(os.environ["MDI_ATTN_CLUSTER"]))
        if os.environ.get("MDI_L2_PF_MB"):
            handle.mdi_set_l2_prefetch_mb(int(os.environ["MDI_L2_PF_MB"]))
        return handle
    except BaseException as e:  # noqa: BLE001
        _load_error = e
        raise OpsError(f"cannot load {_build.LIB}: {e}") from e


def set_prefill_attn_pipe(mode: int) -> None:
    """Prefill attention kernel: 2 = pipelined with two softmax warpgroups (default: validated on B200, 9-12 % faster
    than mode 1 at 2-4k tokens), 1 / True = pipelined with one softmax warpgroup (double-buffered K/V^T tiles and score
    matrix), 0 / False = the simple sequential kernel.  ``MDI_PREFILL_ATTN_PIPE`` sets it at load."""
    lib().mdi_set_prefill_attn_pipe(int(mode))


def set_l2_prefetch_mb(mb: int) -> None:
    """L2 look-ahead of the bulk-copy decode linears: each launch asks the TMA engine to pull up to ``mb`` MiB
    of the weights it will stream next into L2 before it waits for its input (0 = off)."""
    lib().mdi_set_l2_prefetch_mb(int(mb))


def set_linear_variant(v: int) -> None:
    """Default weight-streaming path of the decode linears: 0 = LDG register streaming,
    1 = bulk-copy (TMA engine) ring with 4 stages / 1 CTA per SM, 2 = bulk-copy ring with 2 stages."""
    lib().mdi_set_linear_variant(int(v))


def available() -> bool:
    """True when CUDA is present *and* the library loads.  Raises on a CUDA box whose library is
    broken (no silent eager fallback on the GPU)."""
    if not torch.cuda.is_available():
        return False
    lib()
    return True


def require() -> ctypes.CDLL:
    if not torch.cuda.is_available():
        raise OpsError("the sm_100a kernels need a CUDA device")
    return lib()


def check(code: int, what: str = "") -> None:
    if code != 0:
        msg = {-2: "invalid shape/alignment", -3: "unsupported configuration", -4: "no peer access"}.get(code)
        if msg is None:
            msg = lib().mdi_error_string(code).decode()
        raise OpsError(f"{what or 'kernel launch'} failed: {msg} ({code})")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _bf16(t: Optional[torch.Tensor], name: str) -> None:
    if t is not None and (t.dtype != torch.bfloat16 or not t.is_cuda or not t.is_contiguous()):
        raise OpsError(f"{name}: expected a contiguous CUDA bf16 tensor, got {t.dtype} {t.device}")


def _fp8(t: Optional[torch.Tensor], scale: Optional[torch.Tensor], name: str) -> None:
    if t is None:
        return
    if t.element_size() != 1 or not t.is_cuda or not t.is_contiguous():
        raise OpsError(f"{name}: expected a contiguous CUDA fp8 (1-byte) tensor, got {t.dtype} {t.device}")
    if scale is None or scale.dtype != torch.float32 or scale.shape != (t.shape[0], t.shape[1] // 128) or t.shape[1] % 128:
        raise OpsError(f"{name}: block scales must be fp32 [N, K/128] with K a multiple of 128")


# ---------------------------------------------------------------------------------------------------
def linear_decode(
    W: torch.Tensor, x: torch.Tensor, y: torch.Tensor, ctx: torch.Tensor, *,
    W2: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None, bias2: Optional[torch.Tensor] = None,
    norm_w: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, eps: float = 1e-5,
    unit_offset: bool = False, act: str = "none", x_slot_stride: int = 0, res_slot_stride: int = 0,
    y_slot_stride: int = 0, wait_flag: Optional[int] = None, status: Optional[int] = None, wait_max_cycles: int = 0,
    signal_flag: Optional[int] = None, done_ctr: Optional[int] = None, ctas_per_sm: int = 4, use_pdl: bool = False,
    y_ptr: Optional[int] = None, residual_ptr: Optional[int] = None, x_ptr: Optional[int] = None, variant: int = -1,
    stats: Optional[torch.Tensor] = None, trace: Optional[int] = None,
    wscale: Optional[torch.Tensor] = None, wscale2: Optional[torch.Tensor] = None,
    ctx_early: bool = False, dep_wait: Optional[int] = None,
    dep_signal: Optional[int] = None, dep_ctr: Optional[int] = None, hop_ptr: Optional[int] = None,
    hop_slot_stride: int = 0, prefetch: Optional[Tuple[int, int, int]] = None, l2_pf_chunks: int = 0,
    hop_pre: Optional[Tuple[int, int, int]] = None, norm_b: Optional[torch.Tensor] = None, layer_norm: bool = False) -> None:
    """``y = epilogue(W @ norm?(x))`` for one token.  ``*_ptr`` overrides let the output /
    residual / input live in peer-mapped (other GPU) memory that has no torch tensor.
    With ``wscale`` (fp32 ``[N, K/128]``) ``W`` (and ``W2``/``wscale2``) are fp8-e4m3 block-scaled.

    ``hop_ptr`` (with ``signal_flag``): the fused hop by row copy — ``y`` is a local ``[n_slots, N]`` buffer and the
    last CTA copies the finished row to ``hop_ptr + slot * hop_slot_stride`` (peer memory) before releasing the flag;
    ``hop_pre = (ptr, slot_stride, n)`` puts ``n`` more elements (the residual row ``x`` of a ``[x | h]`` message) in
    front of it.
    ``prefetch = (ptr_a, ptr_b, bytes)``: while waiting for its input the kernel pulls that many bytes of each region
    (the NEXT kernel's weights) into L2; ``l2_pf_chunks``: the same for its own rows beyond the shared-memory ring."""
    if wscale is None:
        _bf16(W, "W"); _bf16(W2, "W2")
    else:
        _fp8(W, wscale, "W"); _fp8(W2, wscale2, "W2")
    N, K = W.shape
    out_fp32 = int(y is not None and y.dtype == torch.float32)
    hist = amax = None
    if stats is not None:  # sampler scratch: histogram at offset 0, packed arg-max right after it
        hist, amax = stats.data_ptr(), stats.data_ptr() + 4096 * 4
    check(lib().mdi_linear_decode(
        ptr(W), ptr(W2), ptr(bias), ptr(bias2), x_ptr if x_ptr is not None else ptr(x), ptr(norm_w),
        residual_ptr if residual_ptr is not None else ptr(residual), y_ptr if y_ptr is not None else ptr(y),
        ptr(ctx), x_slot_stride, res_slot_stride, y_slot_stride, N, K, eps, int(unit_offset), ACT[act], out_fp32,
        wait_flag, status, wait_max_cycles, signal_flag, done_ctr, ctas_per_sm, int(use_pdl) | (2 if ctx_early else 0), variant, hist, amax,
        trace, ptr(wscale), ptr(wscale2), dep_wait, dep_signal, dep_ctr, hop_ptr, hop_slot_stride,
        prefetch[0] if prefetch else None, prefetch[1] if prefetch else None, prefetch[2] if prefetch else 0, l2_pf_chunks,
        hop_pre[0] if hop_pre else None, hop_pre[1] if hop_pre else 0, hop_pre[2] if hop_pre else 0,
        ptr(norm_b), int(layer_norm), stream_ptr()), "linear_decode")


def moe_router(Wg: torch.Tensor, x: torch.Tensor, sel: torch.Tensor, wts: torch.Tensor, ctx: torch.Tensor, *, top: int,
               norm_w: Optional[torch.Tensor] = None, norm_b: Optional[torch.Tensor] = None, layer_norm: bool = False,
               eps: float = 1e-5, unit_offset: bool = False, x_slot_stride: int = 0, wait_flag: Optional[int] = None,
               status: Optional[int] = None, wait_max_cycles: int = 0, use_pdl: bool = False, trace: Optional[int] = None,
               x_ptr: Optional[int] = None) -> None:
    """Router of a mixture-of-experts MLP for one token: ``logits = Wg @ norm(x)`` (bf16), the ``top`` best experts
    and the softmax over their logits -> ``sel`` (int32 ``[top]``) / ``wts`` (fp32 ``[top]``, bf16 values) on the
    device — nothing comes back to the host, the expert passes read both through :func:`moe_linear_decode`."""
    _bf16(Wg, "Wg")
    E, K = Wg.shape
    if sel.dtype != torch.int32 or wts.dtype != torch.float32 or sel.numel() < top or wts.numel() < top:
        raise OpsError("moe_router: sel must be int32 [top], wts fp32 [top]")
    check(lib().mdi_moe_router(ptr(Wg), x_ptr if x_ptr is not None else ptr(x), ptr(norm_w), ptr(norm_b), int(layer_norm),
                               ptr(ctx), x_slot_stride, K, E, top, eps, int(unit_offset), ptr(sel), ptr(wts), wait_flag, status,
                               wait_max_cycles, int(use_pdl), trace, stream_ptr()), "moe_router")


def moe_linear_decode(w_ptrs: torch.Tensor, x: torch.Tensor, y: Optional[torch.Tensor], ctx: torch.Tensor, sel: torch.Tensor,
                      wts: torch.Tensor, k: int, *, N: int, K: int, w2_ptrs: Optional[torch.Tensor] = None,
                      prev: Optional[torch.Tensor] = None, norm_w: Optional[torch.Tensor] = None,
                      norm_b: Optional[torch.Tensor] = None, layer_norm: bool = False, residual: Optional[torch.Tensor] = None,
                      residual_ptr: Optional[int] = None, eps: float = 1e-5, unit_offset: bool = False, act: str = "none",
                      x_slot_stride: int = 0, res_slot_stride: int = 0, y_slot_stride: int = 0, y_ptr: Optional[int] = None,
                      status: Optional[int] = None, signal_flag: Optional[int] = None, done_ctr: Optional[int] = None,
                      hop_ptr: Optional[int] = None, hop_slot_stride: int = 0, ctas_per_sm: int = 3, use_pdl: bool = False,
                      trace: Optional[int] = None, x_ptr: Optional[int] = None, sel_early: bool = False) -> None:
    """One expert pass of a routed token.  ``w_ptrs`` / ``w2_ptrs``: int64 ``[E]`` device tables of the experts' weight
    pointers (``[N, K]`` bf16 each); the kernel reads ``sel[k]`` AFTER its dependency wait and streams only that
    expert.  With ``w2_ptrs``: ``y = act(W1 norm(x)) * (W2 norm(x))``; without: the down pass
    ``y = bf16(wts[k] * bf16(W x)) (+ prev) (+ residual)``, optionally finishing the stage's hop like
    :func:`linear_decode`.  ``sel_early``: the router ran at least two launches before this one, so the expert is
    known before the programmatic-dependency wait and the kernel starts streaming its weights at once."""
    for t, n in ((w_ptrs, "w_ptrs"), (w2_ptrs, "w2_ptrs")):
        if t is not None and (t.dtype != torch.int64 or not t.is_cuda):
            raise OpsError(f"moe_linear_decode: {n} must be an int64 device tensor of pointers")
    check(lib().mdi_moe_linear_decode(
        ptr(w_ptrs), ptr(w2_ptrs), ptr(sel), ptr(wts), k, ptr(prev), x_ptr if x_ptr is not None else ptr(x), ptr(norm_w),
        ptr(norm_b), int(layer_norm), residual_ptr if residual_ptr is not None else ptr(residual),
        y_ptr if y_ptr is not None else ptr(y), ptr(ctx), x_slot_stride, res_slot_stride, y_slot_stride, N, K, eps,
        int(unit_offset), ACT[act], status, signal_flag, done_ctr, hop_ptr, hop_slot_stride, ctas_per_sm, int(use_pdl), trace,
        int(sel_early), stream_ptr()), "moe_linear_decode")


def set_moe_variant(v: int) -> None:
    """Weight path of the expert passes: 0 = register-streamed (LDG, default), otherwise the per-warp bulk-copy ring."""
    lib().mdi_set_moe_variant(v)


def qkv_decode(
    W: torch.Tensor, x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, q_out: torch.Tensor, kv_layer: torch.Tensor,
    ctx: torch.Tensor, *, n_head: int, n_groups: int, head_size: int, rope_n_elem: int, max_seq: int,
    bias: Optional[torch.Tensor] = None, norm_w: Optional[torch.Tensor] = None, eps: float = 1e-5,
    unit_offset: bool = False, x_slot_stride: int = 0, wait_flag: Optional[int] = None, status: Optional[int] = None,
    wait_max_cycles: int = 0, ctas_per_sm: int = 4, use_pdl: bool = False, x_ptr: Optional[int] = None,
    variant: int = -1, trace: Optional[int] = None, wscale: Optional[torch.Tensor] = None, ctx_early: bool = False, dep_wait: Optional[int] = None,
    dep_signal: Optional[int] = None, dep_ctr: Optional[int] = None, norm_b: Optional[torch.Tensor] = None,
    layer_norm: bool = False) -> None:
    if wscale is None:
        _bf16(W, "W")
    else:
        _fp8(W, wscale, "W")
    if cos.dtype != torch.float32 or sin.dtype != torch.float32:
        raise OpsError("rope tables must be fp32")
    check(lib().mdi_qkv_decode(
        ptr(W), ptr(bias), x_ptr if x_ptr is not None else ptr(x), ptr(norm_w), ptr(cos), ptr(sin), ptr(q_out),
        ptr(kv_layer), ptr(ctx), x_slot_stride, W.shape[1], n_head, n_groups, head_size, rope_n_elem, max_seq, eps,
        int(unit_offset), wait_flag, status, wait_max_cycles, ctas_per_sm, int(use_pdl) | (2 if ctx_early else 0), variant, trace, ptr(wscale), dep_wait, dep_signal, dep_ctr,
        ptr(norm_b), int(layer_norm), stream_ptr()),
        "qkv_decode")


def attn_decode(q: torch.Tensor, kv_layer: torch.Tensor, y: torch.Tensor, part: torch.Tensor, tickets: torch.Tensor,
                ctx: torch.Tensor, *, n_head: int, n_groups: int, head_size: int, max_seq: int, n_split: int,
                use_pdl: bool = False, trace: Optional[int] = None, dep_wait: Optional[int] = None,
                dep_signal: Optional[int] = None, dep_ctr: Optional[int] = None, status: Optional[int] = None,
                wait_max_cycles: int = 0) -> None:
    """Split-KV decode attention.  ``part``: fp32 ``[H, n_split, hs+2]`` scratch; ``tickets``: zeroed
    int32 ``[G]`` (self-resetting) used by the last CTA of a group to merge the spans."""
    check(lib().mdi_attn_decode(ptr(q), ptr(kv_layer), ptr(y), ptr(part), ptr(tickets), ptr(ctx), n_head, n_groups,
                                head_size, max_seq, n_split, int(use_pdl), trace, dep_wait, dep_signal, dep_ctr, status, wait_max_cycles,
                                stream_ptr()), "attn_decode")


def embed(wte: torch.Tensor, x: torch.Tensor, ctx: torch.Tensor, *, tokens: Optional[torch.Tensor] = None,
          tok_slot_stride: int = 0, wpe: Optional[torch.Tensor] = None, x_slot_stride: int = 0, scale: float = 1.0,
          use_pdl: bool = False) -> None:
    _bf16(wte, "wte")
    check(lib().mdi_embed(ptr(wte), ptr(wpe), ptr(tokens), tok_slot_stride, ptr(ctx), ptr(x), x_slot_stride,
                          wte.shape[1], scale, int(use_pdl), stream_ptr()), "embed")


def rmsnorm_rows(x: torch.Tensor, w: torch.Tensor, eps: float, unit_offset: bool = False) -> torch.Tensor:
    _bf16(x, "x"); _bf16(w, "w")
    y = torch.empty_like(x)
    C = x.shape[-1]
    check(lib().mdi_rmsnorm_rows(ptr(x), ptr(w), ptr(y), x.numel() // C, C, eps, int(unit_offset), stream_ptr()),
          "rmsnorm_rows")
    return y


def sample(logits: torch.Tensor, tokens: torch.Tensor, ctx: torch.Tensor, *, vocab: int, top_k: Optional[int],
           temperature: float, greedy: bool, seed: int, tok_slot_stride: int, logits_slot_stride: int = 0,
           last_token: Optional[torch.Tensor] = None, use_pdl: bool = False, top_p: float = 1.0) -> None:
    """Stand-alone sampler (one CTA): exact top-k list for k <= 1024 without nucleus crop, else the
    whole-vocabulary sampler (any k, top-p by probability-mass radix selection)."""
    if logits.dtype != torch.float32 or tokens.dtype != torch.int32:
        raise OpsError("sample: logits fp32 and tokens int32 expected")
    check(lib().mdi_sample(ptr(logits), logits_slot_stride, ptr(tokens), tok_slot_stride, ptr(last_token), ptr(ctx),
                           vocab, int(top_k or 0), float(temperature), int(greedy), seed & (2 ** 64 - 1), int(use_pdl),
                           float(top_p), stream_ptr()), "sample")


def gemm(a: torch.Tensor, w: torch.Tensor, *, bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
         out: Optional[torch.Tensor] = None, block_n: int = 128, w2: Optional[torch.Tensor] = None,
         bias2: Optional[torch.Tensor] = None, act: str = "silu_gate", out_ptr: Optional[int] = None,
         signal_flag: Optional[int] = None, done_ctr: Optional[torch.Tensor] = None, ctx: Optional[torch.Tensor] = None,
         status: Optional[torch.Tensor] = None, _knobs: Tuple[int, int, int, int] = (0, 0, 0, 0)) -> Optional[torch.Tensor]:
    """``a [M, K] @ w [N, K]^T (+bias) (+residual)`` on the tcgen05 tensor cores (TMA-fed, TMEM
    accumulator) — the prefill GEMM.  bf16 in/out, fp32 accumulate.

    * ``w2``: gated MLP in one pass — ``act(a w^T + bias) * (a w2^T + bias2)``, two TMEM accumulators
      sharing the A tile.
    * ``out_ptr`` (raw device address, may be a peer mapping) + ``signal_flag``/``done_ctr``/``ctx``:
      the fused prefill hop — the epilogue stores straight into the next stage's input buffer and the
      last CTA releases ``flag[ctx.slot] = ctx.signal`` at system scope.  Returns None."""
    _bf16(a, "a"); _bf16(w, "w"); _bf16(bias, "bias"); _bf16(residual, "residual"); _bf16(w2, "w2"); _bf16(bias2, "bias2")
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K or (w2 is not None and tuple(w2.shape) != tuple(w.shape)):
        raise OpsError(f"gemm: operand shapes differ ({tuple(a.shape)} x {tuple(w.shape)})")
    if out_ptr is None:
        if out is None:
            out = torch.empty(M, N, device=a.device, dtype=torch.bfloat16)
        c_ptr = ptr(out)
    else:
        c_ptr, out = out_ptr, None
    check(lib().mdi_gemm_bf16_ex(ptr(a), ptr(w), ptr(w2), c_ptr, ptr(bias), ptr(bias2), ptr(residual), M, N, K,
                                 ACT[act] if w2 is not None else 0, block_n, signal_flag, ptr(done_ctr), ptr(ctx),
                                 ptr(status), *_knobs, stream_ptr()), "gemm_bf16 (tcgen05)")
    return out


def quantize_rows_fp8(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-token block quantisation of activations for the fp8 GEMM: ``x [M, K]`` bf16 -> (``q [M, K]`` e4m3 as
    uint8, ``scale_t [K/128, M_pad]`` fp32, transposed so that the GEMM's promotion warps read it coalesced)."""
    _bf16(x, "x")
    M, K = x.shape
    if K % 128:
        raise OpsError("quantize_rows_fp8: K must be a multiple of 128")
    m_pad = (M + 127) // 128 * 128
    q = torch.empty(M, K, dtype=torch.uint8, device=x.device)
    scale_t = torch.zeros(K // 128, m_pad, dtype=torch.float32, device=x.device)
    check(lib().mdi_quantize_rows_fp8(ptr(x), ptr(q), ptr(scale_t), M, K, m_pad, stream_ptr()), "quantize_rows_fp8")
    return q, scale_t


def gemm_fp8(a8: torch.Tensor, a_scale_t: torch.Tensor, w8: torch.Tensor, w_scale_t: torch.Tensor, *,
             bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
             w2_8: Optional[torch.Tensor] = None, w2_scale_t: Optional[torch.Tensor] = None, bias2: Optional[torch.Tensor] = None,
             act: str = "silu_gate", out_ptr: Optional[int] = None, signal_flag: Optional[int] = None,
             done_ctr: Optional[torch.Tensor] = None, ctx: Optional[torch.Tensor] = None,
             status: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """Block-scaled fp8 GEMM on ``tcgen05.mma kind::f8f6f4`` (csrc/gemm_fp8_tcgen05.cu): e4m3 operands, one fp32
    scale per 128 K elements per row on both sides (``*_scale_t`` are the TRANSPOSED scale tables ``[K/128, rows]``),
    every K block accumulated in a fresh TMEM partial and folded into fp32 registers with its scales.  bf16 output
    with the same bias / residual / gated-activation / fused-hop epilogues as :func:`gemm`."""
    for t, n in ((a8, "a8"), (w8, "w8"), (w2_8, "w2_8")):
        if t is not None and (t.element_size() != 1 or not t.is_cuda or not t.is_contiguous()):
            raise OpsError(f"gemm_fp8: {n} must be a contiguous CUDA 1-byte (e4m3) tensor")
    _bf16(bias, "bias"); _bf16(residual, "residual"); _bf16(bias2, "bias2")
    M, K = a8.shape
    N = w8.shape[0]
    if w8.shape[1] != K or K % 128 or (w2_8 is not None and tuple(w2_8.shape) != tuple(w8.shape)):
        raise OpsError(f"gemm_fp8: operand shapes differ or K not a multiple of 128 ({tuple(a8.shape)} x {tuple(w8.shape)})")
    nkb = K // 128
    if tuple(w_scale_t.shape) != (nkb, N) or a_scale_t.shape[0] != nkb or a_scale_t.shape[1] < M or a_scale_t.dtype != torch.float32:
        raise OpsError("gemm_fp8: scale tables must be fp32 [K/128, rows] (transposed)")
    if w2_8 is not None and (w2_scale_t is None or tuple(w2_scale_t.shape) != (nkb, N)):
        raise OpsError("gemm_fp8: gated mode needs w2_scale_t [K/128, N]")
    if out_ptr is None:
        if out is None:
            out = torch.empty(M, N, device=a8.device, dtype=torch.bfloat16)
        c_ptr = ptr(out)
    else:
        c_ptr, out = out_ptr, None
    check(lib().mdi_gemm_fp8(ptr(a8), ptr(a_scale_t), a_scale_t.shape[1], ptr(w8), ptr(w_scale_t), ptr(w2_8), ptr(w2_scale_t),
                             c_ptr, ptr(bias), ptr(bias2), ptr(residual), M, N, K, ACT[act] if w2_8 is not None else 0,
                             signal_flag, ptr(done_ctr), ptr(ctx), ptr(status), stream_ptr()), "gemm_fp8 (tcgen05 f8f6f4)")
    return out


def attn_prefill(qkv: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, kv_layer: torch.Tensor, slot: int, *,
                 n_head: int, n_groups: int, head_size: int, rope_n_elem: int,
                 scratch: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> torch.Tensor:
    """Causal attention over a whole prompt (positions ``0..T-1``) on tcgen05: RoPE + split of the fused
    QKV GEMM output ``qkv [T, (H+2G)*hs]``, K/V appended to ``kv_layer[slot]`` (``[n_slots, 2, G, S, hs]``),
    flash attention with S and the per-tile P.V product in TMEM.  Returns ``y [T, H*hs]`` (bf16)."""
    _bf16(qkv, "qkv"); _bf16(kv_layer, "kv_layer")
    if cos.dtype != torch.float32 or sin.dtype != torch.float32:
        raise OpsError("rope tables must be fp32")
    T = qkv.shape[0]
    if qkv.shape[1] != (n_head + 2 * n_groups) * head_size or not qkv.is_contiguous():
        raise OpsError("attn_prefill: qkv must be contiguous [T, (H + 2G) * hs]")
    n_slots, _, G, S, hs = kv_layer.shape
    if G != n_groups or hs != head_size:
        raise OpsError("attn_prefill: KV pool shape does not match the head configuration")
    T_pad = (T + 127) // 128 * 128
    if scratch is None:
        scratch = (torch.zeros(n_head, T_pad, head_size, device=qkv.device, dtype=torch.bfloat16),
                   torch.zeros(n_groups, head_size, T_pad, device=qkv.device, dtype=torch.bfloat16))
    q_s, vt_s = scratch
    if tuple(q_s.shape) != (n_head, T_pad, head_size) or tuple(vt_s.shape) != (n_groups, head_size, T_pad):
        raise OpsError("attn_prefill: scratch tensors have the wrong shape")
    y = torch.empty(T, n_head * head_size, device=qkv.device, dtype=torch.bfloat16)
    check(lib().mdi_attn_prefill(ptr(qkv), ptr(cos), ptr(sin), ptr(kv_layer), ptr(q_s), ptr(vt_s), ptr(y), T, T_pad,
                                 int(slot), n_slots, n_head, n_groups, head_size, rope_n_elem, S, stream_ptr()),
          "attn_prefill (tcgen05)")
    return y


def sample_scratch(device: Any) -> torch.Tensor:
    """Zeroed scratch of the fast sampler: logit-key histogram, packed arg-max, candidate list."""
    return torch.zeros(int(lib().mdi_sample_scratch_bytes()) // 4 + 4, dtype=torch.int32, device=device)


def sample_fast(logits: torch.Tensor, scratch: torch.Tensor, tokens: torch.Tensor, ctx: torch.Tensor, *, vocab: int,
                top_k: Optional[int], temperature: float, greedy: bool, seed: int, tok_slot_stride: int,
                last_token: Optional[torch.Tensor] = None, use_pdl: bool = False, top_p: float = 1.0,
                tok_ts: Optional[torch.Tensor] = None, status: Optional[torch.Tensor] = None) -> None:
    """Sampling from statistics the lm_head epilogue left in ``scratch`` (``linear_decode(stats=scratch)``).
    ``top_p < 1``, ``top_k`` beyond 1024 / None, or an overflowing candidate list take the exact whole-vocabulary
    sampler inside the same launch.  ``tok_ts`` (int64 ``[n_slots, stride]``) receives the device time of every token."""
    if logits.dtype != torch.float32 or tokens.dtype != torch.int32:
        raise OpsError("sample_fast: logits fp32 and tokens int32 expected")
    check(lib().mdi_sample_fast(ptr(logits), ptr(scratch), ptr(tokens), tok_slot_stride, ptr(last_token), ptr(ctx),
                                vocab, int(top_k or 0), float(temperature), int(greedy), seed & (2 ** 64 - 1),
                                int(use_pdl), float(top_p), ptr(tok_ts), tok_ts.shape[1] if tok_ts is not None else 0,
                                ptr(status), stream_ptr()), "sample_fast")


def stamp(dst: torch.Tensor) -> None:
    """``dst[0] = %globaltimer`` on the current stream (int64 tensor): time base of the per-token device timeline."""
    check(lib().mdi_stamp(ptr(dst), stream_ptr()), "stamp")


def advance_step(ctx: torch.Tensor, state: torch.Tensor, pos: torch.Tensor, n_slots: int, is_starter: bool,
                 use_pdl: bool = False) -> None:
    check(lib().mdi_advance_step(ptr(ctx), ptr(state), ptr(pos), n_slots, int(is_starter), int(use_pdl), stream_ptr()),
          "advance_step")


class CudaGraph:
    """A captured sequence of launches on the current stream, replayed from C."""

    def __init__(self) -> None:
        self.exec = c_void_p()
        self.n_nodes = 0
        self._stream: Optional[torch.cuda.Stream] = None

    def __enter__(self) -> "CudaGraph":
        self._stream = torch.cuda.Stream()
        self._stream.wait_stream(torch.cuda.current_stream())
        self._ctx = torch.cuda.stream(self._stream)
        self._ctx.__enter__()
        check(lib().mdi_graph_begin(self._stream.cuda_stream), "graph capture begin")
        return self

    def __exit__(self, exc_type: Any, exc: Any, tb: Any) -> None:
        n = c_int(0)
        code = lib().mdi_graph_end(self._stream.cuda_stream, byref(self.exec), byref(n))
        self._ctx.__exit__(exc_type, exc, tb)
        torch.cuda.current_stream().wait_stream(self._stream)
        if exc_type is None:
            check(code, "graph capture end")
            self.n_nodes = n.value

    def launch(self, times: int = 1) -> None:
        check(lib().mdi_graph_launch(self.exec, stream_ptr(), times), "graph launch")

    def __del__(self) -> None:
        try:
            if self.exec and _lib is not None:
                _lib.mdi_graph_destroy(self.exec)
        except Exception:  # noqa: BLE001
            pass
