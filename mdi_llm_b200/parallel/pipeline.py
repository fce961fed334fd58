"""Device-driven recurrent pipeline over NVLink: the product path.

One process per GPU (``torch.distributed``/NCCL for control-plane collectives only).  Every rank
owns one pipeline stage as a :class:`~.engine.FusedStage`; the ring
``starter → sec0 → … → starter`` is closed with *peer-mapped* buffers (CUDA IPC):

* the last down-projection kernel of stage *i* stores its output row straight into stage
  *i+1*'s ``hidden_in[slot]`` through NVLink and then publishes ``flags[slot]`` with a
  system-scope release (``hop_signal`` in ``csrc/common.cuh``);
* the first kernel of stage *i+1*'s step (QKV projection; ``lm_head`` on the starter for the
  wrap-around hop) prefetches its weights, then acquires that flag and reads the row
  (``hop_wait``) — no socket, no pickle, no NCCL call, no host on the path
  (reference: ``connections.py:325-353`` TX, ``:186-214`` RX, ``gptserver.py:924,1082`` H2D).

The schedule is the reference's recurrent pipeline made static: with FIFO queues and a ring the
order in which samples reach a stage is always ``0,1,…,n-1,0,1,…`` (gptserver.py:864-868,
912-1001), so each stage simply replays "serve slot ``t mod n``" — slot/position/flag sequence
numbers are derived on the device by ``advance_step``.  A sample owns one slot per stage buffer
and is in exactly one place at a time, hence no back-pressure protocol is needed.

``mode="host"`` keeps the same kernels and hops but the host feeds each step descriptor from
pinned memory and reads every sampled token back (the end-to-end/streaming mode).
"""
from __future__ import annotations

import os

import ctypes
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import torch

from .. import ops
from ..models.stage import StageModule
from .engine import FusedStage, HopTarget, RawBuffer
from .scheduler import SamplingParams

__all__ = ["DevicePipeline", "connect_ring_local"]


def _nvtx(name: str):
    """NVTX range around a phase of the pipeline (visible in nsys / ncu timelines)."""
    try:
        return torch.cuda.nvtx.range(name)
    except Exception:  # noqa: BLE001  (torch built without nvtx)
        import contextlib

        return contextlib.nullcontext()


class DevicePipeline:
    def __init__(self, model: StageModule, rank: int, world: int, n_samples: int, max_seq_length: int,
                 sampling: Optional[SamplingParams] = None, max_prompt_len: int = 0, use_pdl: bool = True,
                 ctas_per_sm: int = 4, wait_max_cycles: int = 20_000_000_000, exportable: Optional[bool] = None,
                 hop: str = "p2p", weight_dtype: str = "bf16", free_bf16: bool = False) -> None:
        if hop not in ("p2p", "nccl"):
            raise ValueError("hop must be 'p2p' (fused peer stores + flags) or 'nccl' (send/recv baseline)")
        self.hop = hop if world > 1 else "p2p"
        # kernels of a stage step depend on each other through device flags instead of grid completion
        # (MDI_DEP_FLAGS=1; needs PDL, see common.cuh); off: every kernel waits with griddepcontrol.wait
        self.dep_flags = bool(int(os.environ.get("MDI_DEP_FLAGS", "0"))) and use_pdl
        self.edge_groups: Optional[List[Any]] = None
        self.rank, self.world, self.n = rank, world, n_samples
        self.is_starter, self.is_last = rank == 0, rank == world - 1
        exportable = (world > 1) if exportable is None else exportable
        self.stage = FusedStage(model, n_slots=n_samples, max_seq_length=max_seq_length, sampling=sampling,
                                use_pdl=use_pdl, ctas_per_sm=ctas_per_sm, wait_max_cycles=wait_max_cycles,
                                exportable=exportable, weight_dtype=weight_dtype, free_bf16=free_bf16)
        self.model = model.eval()
        self.device = self.stage.device
        self.C = model.config.n_embd
        self.W_in, self.W_out = self.stage.W_in, self.stage.W_out  # message widths (C, or C + I around a cut MLP)
        if self.hop == "nccl" and (self.W_in != self.C or self.W_out != self.C):
            raise ValueError("the NCCL-hop baseline carries plain hidden states: use a whole- or half-layer partition")
        self.max_prompt_len = int(max_prompt_len or max_seq_length)
        # prefill hop payload (T x C per sample) lands here on secondaries
        self.prefill_raw: Optional[RawBuffer] = None
        self.prefill_in: Optional[torch.Tensor] = None
        if world > 1 and not self.is_starter:
            nbytes = n_samples * self.max_prompt_len * self.W_in * 2
            if exportable:
                self.prefill_raw = RawBuffer(nbytes, self.device)
                self.prefill_in = self.prefill_raw.view(0, (n_samples, self.max_prompt_len, self.W_in), torch.bfloat16)
            else:
                self.prefill_in = torch.zeros(n_samples, self.max_prompt_len, self.W_in, dtype=torch.bfloat16, device=self.device)
        self.next_hop: HopTarget = self.stage.hop_self  # world == 1: the ring closes on myself
        self.next_prefill_ptr: int = 0
        self._opened: List[int] = []
        self.prompt_lens: List[int] = []
        self.round = 1
        self.max_new = 0
        self.n_graph_launches = 0   # decode steps enqueued (one CUDA-graph replay covers `steps_per_graph` of them)
        self.n_kernel_launches = 0  # kernels inside those replays
        # device mode: consecutive steps captured in ONE graph keep their programmatic (PDL) edges, so the first
        # kernel of step t+1 is resident and prefetching its weights while the last kernel of step t drains
        # (a graph boundary is a full stop).  Default: one round per graph, at least 8 steps.
        self.steps_per_graph = int(os.environ.get("MDI_STEPS_PER_GRAPH", "0")) or max(n_samples, 8)

    # ---- ring wiring ------------------------------------------------------------------------------
    def export_handles(self) -> Dict[str, Any]:
        st = self.stage
        return {"rank": self.rank, "hidden": st.raw.handle if st.raw else None, "flag_off": st._flag_off,
                "prefill": self.prefill_raw.handle if self.prefill_raw else None,
                "hidden_ptr": st.hidden_in.data_ptr(), "flag_ptr": st.flags.data_ptr(),
                "prefill_ptr": self.prefill_in.data_ptr() if self.prefill_in is not None else 0,
                "device": self.device.index, "pid": os.getpid()}

    def connect_ipc(self, nxt: Dict[str, Any]) -> None:
        """Open the next stage's exported buffers (other process; a stage living in THIS process — several nodes
        of a test in one interpreter — is addressed directly, an IPC handle cannot be opened by its exporter)."""
        lib = ops.lib()
        if nxt.get("pid") == os.getpid():
            if nxt["device"] != self.device.index:
                ops.check(lib.mdi_enable_peer(self.device.index, nxt["device"]), "enable peer access")
            self.next_hop = HopTarget(nxt["hidden_ptr"], nxt["flag_ptr"])
            self.next_prefill_ptr = nxt["prefill_ptr"]
            return
        with torch.cuda.device(self.device):
            p = ctypes.c_void_p()
            ops.check(lib.mdi_p2p_open(nxt["hidden"], ctypes.byref(p)), "open next hidden_in")
            self._opened.append(int(p.value))
            self.next_hop = HopTarget(int(p.value), int(p.value) + nxt["flag_off"])
            if nxt.get("prefill"):
                q = ctypes.c_void_p()
                ops.check(lib.mdi_p2p_open(nxt["prefill"], ctypes.byref(q)), "open next prefill_in")
                self._opened.append(int(q.value))
                self.next_prefill_ptr = int(q.value)

    def connect_distributed(self, group: Any = None) -> None:
        """All-gather the IPC handles over the control plane and open the next rank's."""
        import torch.distributed as dist

        if self.world == 1:
            return
        if self.hop == "nccl":  # baseline hop: one communicator per ring edge, no peer mapping
            from .transport.nccl_p2p import make_edge_groups

            self.edge_groups = make_edge_groups(self.world)
            dist.barrier(group=group)
            return
        infos: List[Any] = [None] * self.world
        dist.all_gather_object(infos, self.export_handles(), group=group)
        self.connect_ipc(infos[(self.rank + 1) % self.world])
        dist.barrier(group=group)

    def close(self) -> None:
        """Unmap the neighbour's buffers and release this stage's exportable allocations."""
        torch.cuda.synchronize(self.device)
        for p in self._opened:
            ops.lib().mdi_p2p_close(p)
        self._opened.clear()
        self.stage._graphs.clear()
        for raw in (self.prefill_raw, self.stage.raw):
            if raw is not None:
                raw.free()
        self.prefill_raw = self.stage.raw = None

    def set_sampling(self, sampling: SamplingParams) -> None:
        """Change the sampling parameters (starter): they are kernel arguments of the captured step graphs, so the
        graphs are dropped and re-captured on the next launch."""
        if sampling != self.stage.sampling:
            self.stage.sampling = sampling
            self.stage._graphs.clear()

    # ---- generation phases ------------------------------------------------------------------------
    def prepare(self, prompts: Sequence[torch.Tensor], max_new_tokens: int) -> None:
        """Reset per-generation state.  ``prompts`` (token ids) must be known on every rank
        (at least their lengths); call on all ranks, then synchronise + barrier before
        :meth:`prefill`."""
        st = self.stage
        if len(prompts) != self.n:
            raise ValueError(f"pipeline was built for {self.n} samples, got {len(prompts)}")
        self.prompt_lens = [int(p.numel()) for p in prompts]
        if any(t + max_new_tokens > st.S for t in self.prompt_lens):
            raise ValueError(f"Cannot generate {max_new_tokens} tokens - would exceed block size!")
        if max(self.prompt_lens) > self.max_prompt_len and self.world > 1:
            raise ValueError("prompt longer than the prefill hop buffer")
        self.max_new = max_new_tokens
        self.round = 1
        with torch.cuda.device(self.device):
            if not st._graphs:
                st.warmup()
            st.flags.zero_()
            st.status.zero_()
            st.done_ctr.zero_()
            st.reset_deps()
            st.state.copy_(torch.tensor([0, 1, 0, 0], dtype=torch.int32))
            st.pos_arr.copy_(torch.tensor(self.prompt_lens, dtype=torch.int32))
            if self.is_starter:
                st.tokens.zero_()
                st.tok_ts.zero_()
                for i, p in enumerate(prompts):
                    st.tokens[i, : p.numel()].copy_(p.to(torch.int32), non_blocking=True)
            self.prompts = [p.to(self.device, non_blocking=True) for p in prompts] if self.is_starter else None
            if self.hop != "nccl":  # capture (once) every graph the device-driven rounds will replay
                self._g_full(True, self.steps_per_graph)
                self._g_full(True, 1)
                if self.is_starter:
                    self._g_head(True)
            torch.cuda.current_stream().synchronize()

    def _hop_copy(self, src: torch.Tensor, dst_ptr: int) -> None:
        """Prefill hop: peer copy and flag publication fused in one kernel (writers fence at system
        scope before the flag is released — see ``copy_signal_kernel``)."""
        st = self.stage
        ops.check(ops.lib().mdi_copy_signal(src.data_ptr(), dst_ptr, src.numel() * src.element_size(),
                                            self.next_hop.flag_ptr, st.done_ctr.data_ptr(), st.ctx.data_ptr(),
                                            st.status.data_ptr(), ops.stream_ptr()), "prefill hop")

    @torch.inference_mode()
    def prefill(self) -> None:
        """Round 0: every sample's prompt through all stages (tcgen05 GEMMs, see ``FusedStage.prefill``).
        Inter-stage hop: the stage's last down-projection GEMM stores its output tiles straight into
        the next stage's prefill buffer over NVLink and publishes the flag from its last CTA — no copy
        kernel, no NCCL.  The last stage returns only the final row to the starter (8 KB copy+signal)."""
        st, lib = self.stage, ops.lib()
        if self.is_starter:
            with torch.cuda.device(self.device):
                ops.stamp(st.t0_ts)  # time base of the device timeline (the reference's clock also starts before prefill)
        if self.hop == "nccl":
            return self._prefill_nccl()
        with torch.cuda.device(self.device), _nvtx(f"mdi.prefill[{self.rank}]"):
            for slot in range(self.n):
                T = self.prompt_lens[slot]
                pos = torch.arange(T, device=self.device)
                st.set_ctx(slot, T - 1, wait=1, signal=1)
                hop = None if self.is_last else (self.next_prefill_ptr + slot * self.max_prompt_len * self.W_out * 2,
                                                 self.next_hop.flag_ptr)
                if self.is_starter:
                    hidden = st.prefill(self.prompts[slot].view(1, -1), pos, slot, hop=hop)
                else:
                    ops.check(lib.mdi_wait_flag(st.flags.data_ptr(), st.ctx.data_ptr(), st.status.data_ptr(),
                                                st.wait_max_cycles, ops.stream_ptr()), "wait prefill")
                    hidden = st.prefill(self.prefill_in[slot, :T].unsqueeze(0), pos, slot, hop=hop)
                if self.is_last:  # wrap-around: only the last position feeds lm_head
                    self._hop_copy(hidden[0, -1].to(torch.bfloat16).contiguous(), self.next_hop.hidden_ptr + slot * self.C * 2)

    # ---- NCCL-hop baseline ("ours with NCCL send/recv instead of the fused hop") --------------------
    def _edges(self):
        import torch.distributed as dist

        prev, nxt = (self.rank - 1) % self.world, (self.rank + 1) % self.world
        return dist, prev, nxt, self.edge_groups[prev], self.edge_groups[self.rank]

    @torch.inference_mode()
    def _prefill_nccl(self) -> None:
        dist, prev, nxt, g_in, g_out = self._edges()
        st = self.stage
        with torch.cuda.device(self.device):
            for slot in range(self.n):
                T = self.prompt_lens[slot]
                pos = torch.arange(T, device=self.device)
                if self.is_starter:
                    hidden = st.prefill(self.prompts[slot].view(1, -1), pos, slot)
                else:
                    buf = torch.empty(1, T, self.C, dtype=torch.bfloat16, device=self.device)
                    dist.recv(buf, src=prev, group=g_in)
                    hidden = st.prefill(buf, pos, slot)
                hidden = hidden.to(torch.bfloat16).contiguous()
                dist.send(hidden[0, -1].contiguous() if self.is_last else hidden, dst=nxt, group=g_out)

    def _g_nccl(self, head_only: bool) -> ops.CudaGraph:
        st = self.stage

        def build() -> None:
            ops.advance_step(st.ctx, st.state, st.pos_arr, self.n, self.is_starter, use_pdl=False)
            if self.is_starter:
                st.enqueue_head(wait=False)
                st.enqueue_sample()
                if not head_only:
                    st.enqueue_embed(from_tokens=True)
            if not head_only:
                st.enqueue_blocks(None, wait_input=False)

        return st.graph(("nccl", head_only), build, warm=False)

    def _decode_rounds_nccl(self, n_rounds: int) -> int:
        dist, prev, nxt, g_in, g_out = self._edges()
        st = self.stage
        launched = 0
        with torch.cuda.device(self.device):
            for _ in range(n_rounds):
                r = self.round
                if r > self.max_new:
                    break
                final = r == self.max_new
                for slot in range(self.n):
                    if final and not self.is_starter:
                        continue
                    dist.recv(st.hidden_in[slot], src=prev, group=g_in)  # stream-ordered, host does not block
                    g = self._g_nccl(final)
                    g.launch()
                    self.n_kernel_launches += g.n_nodes
                    launched += 1
                    if not final:
                        dist.send(st.out_local[slot], dst=nxt, group=g_out)
                self.round += 1
        self.n_graph_launches += launched
        return launched

    # graph builders ---------------------------------------------------------------------------------
    def _g_full(self, dev_ctx: bool, k: int = 1) -> ops.CudaGraph:
        """``k`` consecutive full steps (device mode; host-fed steps are one per graph)."""
        st = self.stage

        def build() -> None:
            for j in range(k):
                if dev_ctx:
                    ops.advance_step(st.ctx, st.state, st.pos_arr, self.n, self.is_starter, use_pdl=self.stage.use_pdl and j > 0)
                if self.is_starter:
                    st.enqueue_head(wait=True)
                    st.enqueue_sample()
                    st.enqueue_embed(from_tokens=True)
                st.enqueue_blocks(self.next_hop, wait_input=True, dep_flags=self.dep_flags)

        return st.graph(("full", dev_ctx, self.next_hop.hidden_ptr, self.dep_flags, k), build, warm=False)

    def _g_head(self, dev_ctx: bool) -> ops.CudaGraph:
        st = self.stage

        def build() -> None:
            if dev_ctx:
                ops.advance_step(st.ctx, st.state, st.pos_arr, self.n, True, use_pdl=False)
            st.enqueue_head(wait=True)
            st.enqueue_sample()

        return st.graph(("head", dev_ctx), build, warm=False)

    def decode_rounds(self, n_rounds: int) -> int:
        """Enqueue ``n_rounds`` decode rounds (every sample advances one token per round) in
        device-driven mode.  Returns the number of steps issued."""
        st = self.stage
        if self.hop == "nccl":
            return self._decode_rounds_nccl(n_rounds)
        launched = 0
        with torch.cuda.device(self.device), _nvtx(f"mdi.decode[{self.rank}] x{n_rounds}"):
            full_rounds = max(0, min(n_rounds, self.max_new - self.round))  # rounds before the final (head-only) one
            steps = full_rounds * self.n
            # only two graph shapes ever exist (k steps, 1 step), both captured in prepare(): a segment never pays
            # for a capture inside somebody's timed region
            k = self.steps_per_graph if steps >= self.steps_per_graph else 1
            if steps:
                g = self._g_full(True, k)
                g.launch(steps // k)
                self.n_kernel_launches += g.n_nodes * (steps // k)
                if steps % k:
                    g1 = self._g_full(True, 1)
                    g1.launch(steps % k)
                    self.n_kernel_launches += g1.n_nodes * (steps % k)
                launched += steps
                self.round += full_rounds
            if n_rounds > full_rounds and self.round == self.max_new:
                if self.is_starter:
                    g = self._g_head(True)
                    g.launch(self.n)
                    self.n_kernel_launches += g.n_nodes * self.n
                    launched += self.n
                self.round += 1
        st._step_seq += launched  # advance_step numbered these steps on the device
        self.n_graph_launches += launched
        return launched

    def decode_rounds_host(self, n_rounds: int, on_token: Optional[Callable[[int, int, int], None]] = None) -> Tuple[int, int, int]:
        """Host-fed variant (end-to-end / streaming mode): per step the descriptor goes H2D from pinned memory and,
        on the starter, the sampled token comes back D2H into pinned memory and is handed to ``on_token``.  The read-back
        of step t is awaited AFTER step t+1 has been issued (its descriptor does not depend on the token — the device
        reads tokens from its own ring), so the GPU is not idle during the host's turnaround.
        Returns (launches, h2d_bytes, d2h_bytes)."""
        st = self.stage
        launched = h2d = d2h = 0
        use_events = self.is_starter and st.last_token.is_cuda
        if use_events and not hasattr(self, "_tok_pinned"):
            self._tok_pinned = torch.zeros(4, dtype=torch.int32).pin_memory()
            self._tok_events = [torch.cuda.Event() for _ in range(4)]
        pending: Optional[Tuple[int, int, int]] = None  # (ring index, slot, pos) of the step whose token is still in flight

        def deliver(p: Tuple[int, int, int]) -> None:
            i, slot_, pos_ = p
            self._tok_events[i].synchronize()
            if on_token is not None:
                on_token(slot_, pos_, int(self._tok_pinned[i]))

        with torch.cuda.device(self.device):
            for _ in range(n_rounds):
                r = self.round
                if r > self.max_new:
                    break
                final = r == self.max_new
                for slot in range(self.n):
                    pos = self.prompt_lens[slot] + r - 1
                    st.set_ctx(slot, pos, wait=r if self.is_starter else r + 1, signal=r + 1)
                    h2d += ops.CTX_INTS * 4
                    if final and not self.is_starter:
                        continue
                    g = self._g_head(False) if final else self._g_full(False)
                    g.launch()
                    self.n_kernel_launches += g.n_nodes
                    launched += 1
                    if not self.is_starter and launched % 2048 == 0:
                        torch.cuda.current_stream().synchronize()  # keep the pinned ctx ring ahead of the GPU
                    if self.is_starter:
                        d2h += 4
                        if use_events:
                            i = launched % 4
                            self._tok_pinned[i: i + 1].copy_(st.last_token[slot: slot + 1], non_blocking=True)
                            self._tok_events[i].record()
                            if pending is not None:
                                deliver(pending)
                            pending = (i, slot, pos)
                        else:  # CPU dry run of the orchestration
                            tok = int(st.last_token[slot].item())
                            if on_token is not None:
                                on_token(slot, pos, tok)
                self.round += 1
            if pending is not None:
                deliver(pending)
        self.n_graph_launches += launched
        return launched, h2d, d2h

    def result_tokens(self) -> Dict[int, torch.Tensor]:
        """starter: prompt + generated ids per sample (syncs)."""
        st = self.stage
        assert self.is_starter
        torch.cuda.synchronize(self.device)
        err, aborted = (int(x) for x in st.status[:2].tolist())
        if err & 3 or aborted:
            raise RuntimeError("pipeline aborted: " + ("hop watchdog expired on this stage (a neighbour stopped responding)"
                                                       if err & 1 else "abort flag received from the ring"))
        done = min(self.round - 1, self.max_new)
        out = {}
        for i, T in enumerate(self.prompt_lens):
            out[i] = st.tokens[i, : T + done].to("cpu", torch.int64).view(1, -1)
        return out

    def token_times(self) -> List[float]:
        """starter: seconds (device clock, since the start of prefill) at which the 1st, 2nd, ... generated token
        of the run was sampled — the reference's ``tok_time`` (gptserver.py:952-956) without a host in the loop."""
        st = self.stage
        assert self.is_starter
        torch.cuda.synchronize(self.device)
        done = min(self.round - 1, self.max_new)
        t0 = int(st.t0_ts.item())
        ts: List[int] = []
        for i, T in enumerate(self.prompt_lens):
            ts += st.tok_ts[i, T: T + done].tolist()
        return [(t - t0) * 1e-9 for t in sorted(x for x in ts if x > 0)]

    def poison(self) -> None:
        """Abort from the host: overwrite this stage's incoming flags with the poison value on a side stream.
        Queued steps stop waiting, mark the stage aborted and publish poison downstream (common.cuh)."""
        side = torch.cuda.Stream(device=self.device)
        with torch.cuda.stream(side):
            self.stage.flags.fill_(ops.POISON)
        side.synchronize()

    def generate(self, prompts: Sequence[torch.Tensor], max_new_tokens: int, sync: Optional[Callable[[], None]] = None,
                 mode: str = "device") -> Optional[Dict[int, torch.Tensor]]:
        """Whole generation on this rank.  ``sync`` = cross-rank barrier (None when world == 1)."""
        self.prepare(prompts, max_new_tokens)
        if sync is not None:
            sync()
        self.prefill()
        if mode == "device":
            self.decode_rounds(max_new_tokens)
        else:
            self.decode_rounds_host(max_new_tokens)
        torch.cuda.synchronize(self.device)
        out = self.result_tokens() if self.is_starter else None
        if sync is not None:
            sync()
        return out


def connect_ring_local(pipes: Sequence[DevicePipeline]) -> None:
    """Wire pipelines that live in ONE process (one per GPU) with direct peer access."""
    n = len(pipes)
    lib = ops.lib()
    for i, p in enumerate(pipes):
        nxt = pipes[(i + 1) % n]
        if p.device != nxt.device:
            ops.check(lib.mdi_enable_peer(p.device.index, nxt.device.index), "enable peer access")
        p.next_hop = HopTarget(nxt.stage.hidden_in.data_ptr(), nxt.stage.flags.data_ptr())
        p.next_prefill_ptr = nxt.prefill_in.data_ptr() if nxt.prefill_in is not None else 0
