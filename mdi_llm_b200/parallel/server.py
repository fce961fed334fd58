"""Node runtime: one pipeline stage + its control endpoint + its place in the ring.

Parity: reference ``src/sub/gptserver.py`` ``GPTServer`` — constructor contract (:139-325),
``launch_starter`` (:358-394), ``start_inference`` (:396-474), ``stop_generation`` (:476-493),
``shutdown`` (:495-514), device selection priority CLI > JSON ``device`` > default (:601-617),
meta-device model construction + chunk loading with dtype override (:619-714), tokenizer and
prompt-style loading (:716-749), REST verbs ``GET /``, ``POST /init``, ``PUT /stop``, ``DELETE``
(:1114-1226).  The generation loops themselves live in :mod:`.scheduler`.

State is per instance (the reference keeps it in class attributes, one node per process,
gptserver.py:72-137), so several nodes can share a process — which is what the CPU tests do.

Data plane (``transport=``): ``"socket"`` is the reference's TCP+pickle ring driven by the host loops of
:mod:`.scheduler`; ``"p2p"`` (what ``"auto"`` picks when every node is a CUDA device of this box and the
architecture is covered by the fused kernels) is the device ring of :mod:`.ring` / :mod:`.pipeline`:
hop buffers exported as CUDA-IPC handles in the ``POST /init`` response, neighbours mapped at
``POST /ring {"op": "connect"}``, generation enqueued as CUDA-graph replays with the hop fused into the
kernels; ``"nccl"`` keeps those kernels but moves the hop to NCCL send/recv (the comparison midpoint).
"""
from __future__ import annotations

import gc
import json
import logging
import threading
import time
import warnings
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import torch

from .. import config as C
from ..models.config import Config
from ..models.partition import count_transformer_blocks, plan_layers, stage_shape_from_state_dict
from ..models.stage import StageModule, build_stage
from ..text.prompts import PromptStyle, get_user_prompt, has_prompt_style, load_prompt_style
from ..text.tokenizer import Tokenizer
from ..utils.checkpoint import lazy_load, materialize_stage
from ..utils.context_managers import catch_loop_errors
from ..utils.misc import find_eot, waiting_animation
from ..utils.safe_pickle import safe_loads
from .control import ControlServer, HTTPError, is_loopback
from .scheduler import (EagerStageRunner, GenerationResult, SamplingParams, StageRunner, secondary_loop,
                        starter_loop)
from .transport import LoopbackTransport, SocketTransport, Transport

logger_wp = logging.getLogger("model_dist")
logger_wp.setLevel(logging.ERROR)

FileType = Union[str, Path]
__all__ = ["GPTServer"]


def _resolve_dtype(dtype: Optional[str], device: str) -> Tuple[str, torch.dtype]:
    name = dtype if dtype else (C.default_dtype() if "cuda" in device else "float32")
    if name not in C.DTYPE_TORCH_MAPPING:
        raise ValueError(f"Unsupported dtype {name!r}; choose from {list(C.DTYPE_TORCH_MAPPING)}")
    if name == "bfloat16" and "cuda" in device and not (torch.cuda.is_available() and torch.cuda.is_bf16_supported()):
        raise ValueError("Specified bfloat16, but the host does not support this format")
    return name, C.DTYPE_TORCH_MAPPING[name]


class GPTServer:
    def __init__(
        self,
        node_config: Dict[str, Any],
        node_type: str,
        *,
        model_config: Optional[Config] = None,
        chunk_path: Optional[FileType] = None,
        tokenizer_dir: Optional[FileType] = None,
        model_device: Optional[str] = None,
        dtype: Optional[str] = None,
        **kwargs: Any,
    ) -> None:
        self.verb = bool(kwargs.get("verb", False))
        self.plots = bool(kwargs.get("plots", False))
        self.model_type: Optional[str] = kwargs.get("model_type")
        self.max_seq_length: Optional[int] = kwargs.get("model_seq_length")
        self.compile = bool(kwargs.get("compile", False))
        self.engine_kind: str = kwargs.get("engine", "auto")  # "auto" | "eager" | "cuda"
        self.transport_kind: str = kwargs.get("transport", "auto")  # auto | socket | p2p | nccl
        if self.transport_kind not in ("auto", "socket", "p2p", "nccl"):
            raise ValueError(f"transport must be auto, socket, p2p or nccl, got {self.transport_kind!r}")
        self.weights: str = kwargs.get("weights", "bf16")          # bf16 | fp8 (block-scaled, device ring only)
        self.decode_mode: str = kwargs.get("decode_mode", "device")  # device ring: device-driven | host-fed steps
        self.random_init: Optional[int] = kwargs.get("random_init")  # seed: synthetic weights instead of a chunk
        self.stage_spec: Optional[Dict[str, Any]] = kwargs.get("stage_spec")
        self.max_prompt_len: int = int(kwargs.get("max_prompt_len") or 0)
        self.on_token = kwargs.get("on_token")  # streaming callback(sample, position, token) in host-fed mode
        self.token = kwargs.get("token")
        self.insecure = kwargs.get("insecure")
        self.ring: Optional[Any] = None  # RingBackend when the device ring is the data plane
        self.ring_transport: Optional[str] = None
        self.peer_handles: List[Dict[str, Any]] = []
        self.last_ring_stats: Optional[Dict[str, Any]] = None
        self.sampling: SamplingParams = kwargs.get("sampling") or SamplingParams(
            temperature=kwargs.get("temperature", C.TEMPERATURE), top_k=kwargs.get("top_k", C.TOP_K),
            top_p=kwargs.get("top_p", 1.0), seed=kwargs.get("seed"))
        self.chaos = kwargs.get("chaos")
        self.watchdog_s: Optional[float] = kwargs.get("watchdog_s")
        self.start_http = bool(kwargs.get("start_http", True))
        # earlier protocol generations: cache-less context re-send; head on the last node
        self.use_kv_cache = bool(kwargs.get("use_kv_cache", True))
        self.head_on: str = kwargs.get("head_on", "starter")
        self.requested_dtype = dtype

        # --- per-instance run-time state ----------------------------------------------------
        self.model: Optional[StageModule] = None
        self.runner: Optional[StageRunner] = None
        self.transport: Optional[Transport] = None
        self.running = threading.Event()
        self.inference_thread: Optional[threading.Thread] = None
        self.webserv: Optional[ControlServer] = None
        self.tok: Optional[Tokenizer] = None
        self.prompt_style: Optional[PromptStyle] = None
        self.stop_tokens: Tuple[List[int], ...] = ()
        self.tok_time: List[Tuple[int, float]] = []
        self.last_result: Optional[GenerationResult] = None
        self.n_samples: Optional[int] = None
        self.prev_node: Optional[Dict[str, Any]] = None
        self.next_node: Optional[Dict[str, Any]] = None
        self.loop_error: Optional[BaseException] = None
        self._initializing = threading.Lock()

        self.node_type = node_type
        self.node_config = node_config

        if "starter" in node_type:
            assert chunk_path is not None, "Missing path to the model chunk"
            assert model_config is not None, "Missing model Config"
            assert tokenizer_dir is not None, "Missing tokenizer directory"
            self.model_path = Path(chunk_path)
            self.tokenizer_dir = Path(tokenizer_dir)
            if self.model_type is None:
                self.model_type = self.tokenizer_dir.name or None
            if self.plots and self.model_type is None:
                raise ValueError("-p flag requires to correctly set the model type")
            self.role = "starter"
            self.own_config = node_config["nodes"]["starter"]
            self._select_device(model_device)
            secondaries = node_config["nodes"].get("secondary", [])
            self.n_nodes = 1 + len(secondaries)
            self.next_node = None if self.n_nodes == 1 else secondaries[0]
            self.prev_node = None if self.n_nodes == 1 else secondaries[-1]
            self.model_config = model_config
            self.n_layers_local = self._infer_local_layers(self.model_path, "starter")
            self._init_model(self.n_layers_local, model_path=None if self.random_init is not None else self.model_path)
            self._load_tokenizer(self.tokenizer_dir)
        else:
            self.model_config = model_config
            self.model_path = Path(chunk_path) if chunk_path is not None else None
            parts = node_type.split(":")
            if len(parts) == 1:
                secs = node_config.get("nodes", {}).get("secondary", []) if "nodes" in node_config else [node_config]
                if len(secs) != 1:
                    raise ValueError("Need to specify which of the secondary nodes this is ('secondary:n')")
                secondary_index = 0
            else:
                secondary_index = int(parts[1])
            self.role = self.node_type = f"secondary:{secondary_index}"
            self.own_config = node_config if "nodes" not in node_config else node_config["nodes"]["secondary"][secondary_index]
            self.starter_addr = self.own_config["communication"].get("starter_addr")
            self._select_device(model_device)
            self.n_nodes = None

        self.own_addr = self.own_config["addr"]
        self.own_comm_port = self.own_config["communication"]["port"]
        self.inference_port_in = self.own_config["inference"]["port_in"]
        self.inference_port_out = self.own_config["inference"]["port_out"]
        if self.start_http:
            self.start_webserv()

    # ---- web server ---------------------------------------------------------------------------
    def start_webserv(self) -> None:
        self.webserv = ControlServer(self, self.own_addr, self.own_comm_port, token=self.token, insecure=self.insecure)
        self.webserv.start()

    def stop_webserv(self) -> None:
        if self.webserv is not None:
            self.webserv.stop()

    def block(self) -> None:
        if self.webserv is not None:
            self.webserv.block()

    # ---- model / tokenizer ------------------------------------------------------------------------
    def _select_device(self, device: Optional[str]) -> None:
        if device:
            self.model_device = device
        elif "device" in self.own_config:
            self.model_device = self.own_config["device"]
        else:
            warnings.warn(f"Using default device {C.DEVICE}")
            self.model_device = C.DEVICE
        self.torch_model_device = torch.device(self.model_device)
        self.dtype, self.ptdtype = _resolve_dtype(self.requested_dtype, self.model_device)
        if self.verb:
            print(f"Using device: {self.model_device}, dtype {self.dtype}")

    def _infer_local_layers(self, chunk: Optional[Path], role: str) -> int:
        """Layer count of this node: read it off the chunk file when there is one (works for
        any partition plan), else fall back to the planner / reference table."""
        assert self.model_config is not None
        if self.random_init is not None and self.stage_spec:
            return int(self.stage_spec["n_blocks"])
        if chunk is not None and Path(chunk).is_file():
            n = count_transformer_blocks(lazy_load(chunk))
            if n:
                return n
        if self.n_nodes in (None, 1):
            return self.model_config.n_layer
        plan = plan_layers(self.n_nodes, self.model_config.n_layer, self.model_config)
        return plan[0] if role == "starter" else plan[1 + int(role.split(":")[1])]

    def _init_model(self, n_transf_layers: int, *, model_path: Optional[Path] = None,
                    model_parameters: Optional[Dict[str, Any]] = None) -> None:
        assert self.model_config is not None, "No model configuration was found!"
        assert self.model is None, "The model was already initialized!"
        if not (model_path or model_parameters) and self.random_init is None:
            raise ValueError("At least one between model_path and model_parameters must be nonempty")
        role, extra = self.node_type, {}
        if self.head_on == "finisher" and (self.n_nodes or 1) > 1:
            if "starter" in role:
                extra["with_head"] = False
            elif int(role.split(":")[1]) == self.n_nodes - 2:
                role = "finisher"
        sd = None
        if model_path or model_parameters is not None:
            sd = lazy_load(model_path) if model_path else model_parameters
            shape = stage_shape_from_state_dict(sd)  # half-layer chunks describe themselves
        else:
            shape = dict(self.stage_spec or {})
        for k, legacy, val in (("first_parts", "first_mlp_only", "mlp"), ("last_parts", "last_attn_only", "attn")):
            parts = shape.get(k) or (val if shape.get(legacy) else "both")
            if parts != "both":
                extra[k] = parts
        model = build_stage(self.model_config, role, n_transf_layers, meta=True, verb=self.verb, **extra)
        if sd is None:  # synthetic weights (benchmarks on a box without checkpoints): same model for any partition
            from ..utils.checkpoint import random_init_stage_

            random_init_stage_(model, self.torch_model_device, self.ptdtype, seed=int(self.random_init),
                               layer_offset=int((self.stage_spec or {}).get("layer_offset", 0)))
        else:
            wanted = {k for k, _ in model.named_parameters()}
            if model_path and "starter" in self.node_type and self.n_nodes == 1:
                # standalone: the chunk is the full lit_model.pth — keys already match the starter
                sd = {k: v for k, v in sd.items() if k in wanted or k == "lm_head.weight"}
            materialize_stage(model, dict(sd), self.torch_model_device, self.ptdtype)
        if self.max_seq_length:
            model.max_seq_length = self.max_seq_length
            model.cos, model.sin = model.cos.to(self.torch_model_device), model.sin.to(self.torch_model_device)
        else:
            self.max_seq_length = model.max_seq_length
        self.model = model.eval()
        self.runner = None  # built on first use: sized from n_samples, and not at all when the device ring runs
        del sd
        gc.collect()

    def _make_runner(self, model: StageModule, n_samples: Optional[int] = None) -> StageRunner:
        kind = self.engine_kind
        legacy = not self.use_kv_cache or (self.head_on == "finisher" and (self.n_nodes or 1) > 1)
        if kind in ("auto", "cuda") and self.torch_model_device.type == "cuda" and not legacy:
            try:
                from .engine import FusedStageRunner, engine_supports

                if engine_supports(model.config, self.ptdtype):
                    return FusedStageRunner(model, max_seq_length=model.max_seq_length, n_slots=max(1, int(n_samples or 8)),
                                            sampling=self.sampling)
                if kind == "cuda":
                    raise RuntimeError(f"CUDA engine does not support config {model.config.name}")
                self._hint_fit_engine(model.config)
            except ImportError:
                if kind == "cuda":
                    raise
        return EagerStageRunner(model, n_slots_hint=max(1, int(n_samples or 1)))

    @staticmethod
    def _hint_fit_engine(config: Config) -> None:
        """The model runs on the eager fallback although an exact re-parametrisation would put it on the fused kernels."""
        try:
            from ..utils.fit_engine import fit_engine

            _, _, notes = fit_engine(config)
        except Exception:  # noqa: BLE001  (outside the engine for another reason: nothing to suggest)
            return
        if notes:
            warnings.warn(f"{config.name} runs on the eager PyTorch fallback ({'; '.join(notes)} would be needed): "
                          "`python -m mdi_llm_b200.cli.prepare_model <checkpoint> --fit-engine` writes an exactly equivalent "
                          "checkpoint that the fused sm_100a engine covers")

    # ---- device ring (transport p2p / nccl) --------------------------------------------------------------
    def resolve_transport(self) -> str:
        """The data plane this run will use.  ``auto`` → ``p2p`` iff this node can run the fused ring and
        every node of the topology lives on this box (CUDA IPC cannot cross hosts) on a CUDA device."""
        legacy = not self.use_kv_cache or self.head_on == "finisher"
        kind = self.transport_kind
        if kind == "socket" or self.engine_kind == "eager" or legacy or self.chaos is not None:
            if kind in ("p2p", "nccl"):
                raise ValueError(f"transport {kind!r} needs the fused engine with KV caches (no chaos policy, no legacy protocol)")
            return "socket"
        from .ring import ring_capable

        assert self.model_config is not None
        capable = ring_capable(self.model_device, self.ptdtype, self.model_config)
        if kind in ("p2p", "nccl"):
            if not capable:
                raise ValueError(f"transport {kind!r}: {self.model_config.name} / {self.dtype} on {self.model_device} is not "
                                 "covered by the fused engine — use --transport socket")
            return kind
        nodes = self.node_config.get("nodes", {}) if "nodes" in self.node_config else {}
        every = [nodes.get("starter", self.own_config)] + list(nodes.get("secondary", []))
        one_box = all(is_loopback(str(n["addr"])) for n in every) or len({str(n["addr"]) for n in every}) == 1
        all_cuda = all(str(n.get("device", self.model_device)).startswith("cuda") for n in every)
        return "p2p" if capable and one_box and all_cuda else "socket"

    def ring_setup(self, n_samples: int, rank: int, world: int, transport: str = "p2p") -> Dict[str, Any]:
        """Build this node's :class:`RingBackend` (hop buffers allocated exportable) and return the CUDA-IPC
        handles a predecessor needs to store into them."""
        from .ring import RingBackend

        assert self.model is not None
        if self.ring is not None and (self.ring.n_samples != n_samples or self.ring.world != world):
            self.ring.close()
            self.ring = None
        if self.ring is None:
            cycles = int(self.watchdog_s * 1.9e9) if self.watchdog_s else 20_000_000_000
            self.ring = RingBackend(self.model, rank, world, n_samples, self.model.max_seq_length, sampling=self.sampling,
                                    max_prompt_len=self.max_prompt_len or self.model.max_seq_length,
                                    hop="nccl" if transport == "nccl" else "p2p", weight_dtype=self.weights,
                                    wait_max_cycles=cycles)
        self.ring_transport = transport
        return self.ring.handles()

    def _load_tokenizer(self, tokenizer_dir: FileType) -> None:
        d = Path(tokenizer_dir)
        try:
            try:
                self.tok = Tokenizer(d, force_backend="huggingface")
            except Exception:  # noqa: BLE001  (broken tokenizer_config.json, missing file ...)
                self.tok = Tokenizer(d)
        except (NotImplementedError, FileNotFoundError):
            warnings.warn(f"No tokenizer files in {d}: using the byte-level tokenizer")
            from ..text.tokenizer import write_bytes_tokenizer

            write_bytes_tokenizer(d)
            self.tok = Tokenizer(d, force_backend="bytes")
        assert self.model_config is not None
        self.prompt_style = load_prompt_style(d) if has_prompt_style(d) else PromptStyle.from_config(self.model_config)
        try:
            self.stop_tokens = self.prompt_style.stop_tokens(self.tok)
        except ValueError:
            self.stop_tokens = ([self.tok.eos_id],)

    # ---- generation -------------------------------------------------------------------------------
    def launch_starter(self, n_samples: int, max_tokens: int,
                       prompt: Optional[Union[str, Sequence[torch.Tensor]]] = None) -> Tuple[List[str], List[Tuple[int, float]]]:
        if self.role != "starter":
            raise ValueError(f"Cannot run `launch_starter` for node type {self.role}")
        metrics: Dict[str, Any] = {}
        self.n_samples = n_samples
        if self.ring is None and (self.n_nodes or 1) == 1 and self.resolve_transport() != "socket":
            self.ring_setup(n_samples, 0, 1, "p2p")  # standalone on a GPU: the ring closes on this node
            self.ring.connect(None)
        self.inference_thread = threading.Thread(
            target=self.start_inference, args=(n_samples,),
            kwargs={"max_new_tokens": max_tokens, "prompt": prompt, "metrics": metrics})
        self.inference_thread.start()
        self.inference_thread.join()
        self.shutdown()
        if self.loop_error is not None:
            raise self.loop_error
        return metrics["gen_text"], metrics["gen_time"]

    def _create_transport(self) -> Transport:
        if self.n_nodes == 1 or (self.role == "starter" and self.next_node is None):
            return LoopbackTransport(chaos=self.chaos)
        if self.prev_node is None or self.next_node is None:
            raise RuntimeError("Missing neighboring node info!")
        return SocketTransport(self.own_config, self.prev_node, self.next_node,
                               is_starter=self.role == "starter", chaos=self.chaos, verb=self.verb)

    def start_inference(self, n_samples: int, *, max_new_tokens: Optional[int] = None,
                        prompt: Optional[Union[str, Sequence[torch.Tensor]]] = None,
                        metrics: Optional[Dict[str, Any]] = None) -> None:
        assert self.model_config is not None and self.model is not None
        if self.ring is not None and self.role == "starter":
            try:
                self.running.set()
                assert max_new_tokens is not None
                out_text, gen_time = self._starter_ring(n_samples, prompt, int(max_new_tokens))
                if metrics is not None:
                    metrics["gen_text"], metrics["gen_time"] = out_text, gen_time
            except BaseException as e:  # noqa: BLE001
                self.loop_error = e
                logger_wp.error(f"device ring failed: {e!r}")
                if metrics is not None:
                    metrics.setdefault("gen_text", [])
                    metrics.setdefault("gen_time", [])
            finally:
                self.running.clear()
            return
        if self.runner is None:
            self.runner = self._make_runner(self.model, n_samples)
        try:
            if self.transport is not None:
                self.transport.shutdown()
            self.transport = self._create_transport()
            self.running.set()
            self.transport.launch()
            logger_wp.info("Starting generation loop")
            if self.role == "starter":
                assert max_new_tokens is not None
                out_text, gen_time = self._starter_loop(n_samples, prompt, max_new_tokens=max_new_tokens)
                if metrics is not None:
                    metrics["gen_text"], metrics["gen_time"] = out_text, gen_time
            else:
                self._secondary_loop()
        except BaseException as e:  # noqa: BLE001  — surfaced by launch_starter / logged on workers
            self.loop_error = e
            self.running.clear()
            logger_wp.error(f"inference loop failed: {e!r}")
            if metrics is not None:
                metrics.setdefault("gen_text", [])
                metrics.setdefault("gen_time", [])

    def encode_prompts(self, prompt: Optional[Union[str, Sequence[torch.Tensor]]], n_samples: int) -> List[torch.Tensor]:
        """Styled + tokenised prompts; a sequence of id tensors is passed through (synthetic runs)."""
        if prompt is not None and not isinstance(prompt, str):
            ids = [torch.as_tensor(p, dtype=torch.int).reshape(-1) for p in prompt]
            if len(ids) != n_samples:
                raise ValueError(f"{len(ids)} token prompts for {n_samples} samples")
            return ids
        assert self.tok is not None and self.prompt_style is not None
        if prompt is None:
            texts = [self.prompt_style.apply("\n") for _ in range(n_samples)]
        else:
            texts = get_user_prompt(prompt, n_samples, prompt_style=self.prompt_style)
        return [self.tok.encode(t) for t in texts]

    def _starter_loop(self, n_samples: int, prompt: Any = None, **kwargs: Any) -> Tuple[List[str], List[Tuple[int, float]]]:
        assert self.model is not None and self.runner is not None and self.transport is not None
        idx = [p.to(self.torch_model_device) for p in self.encode_prompts(prompt, n_samples)]
        S = self.model.max_seq_length
        if "max_new_tokens" in kwargs and kwargs["max_new_tokens"] is not None:
            max_new = int(kwargs["max_new_tokens"])
            if self.use_kv_cache and any(max_new + p.numel() > S for p in idx):  # cache-less mode crops
                raise ValueError(f"Cannot generate {max_new} tokens - would exceed block size!")
        else:
            max_new = S - max(p.numel() for p in idx)
            assert max_new > 0, "Some prompt is longer than the context length of the model"
        spinner_stop = threading.Event()
        spinner = None
        if self.verb:
            spinner = threading.Thread(target=waiting_animation, args=("Processing samples", spinner_stop), daemon=True)
            spinner.start()
        with catch_loop_errors(running_event=self.running, event_to_be_set=[spinner_stop]):
            res = starter_loop(self.runner, self.transport, idx, max_new, self.sampling, self.running,
                               n_nodes=self.n_nodes or 1, record_times=True, watchdog_s=self.watchdog_s,
                               use_kv_cache=self.use_kv_cache, block_size=self.model.max_seq_length,
                               head_remote=self.head_on == "finisher" and (self.n_nodes or 1) > 1)
        self.running.clear()
        self.last_result = res
        self.tok_time = res.tok_time
        logger_wp.info("Generation completed")
        truncated = [find_eot(res.samples[i], self.stop_tokens, res.prompt_lengths[i]) for i in sorted(res.samples)]
        assert self.tok is not None
        return [self.tok.decode(s) for s in truncated], self.tok_time

    def open_ring_session(self, n_samples: int, prompt: Any, max_new_tokens: int, mode: Optional[str] = None,
                          sampling: Optional[SamplingParams] = None) -> Any:
        """Prepared generation over the device ring (prompts encoded, every node reset); drive it with
        ``session.run(rounds)``.  ``launch_starter`` is ``open_ring_session`` + ``run()`` + decode."""
        from .ring import RingSession

        assert self.ring is not None and self.model is not None
        if n_samples < 1:  # the same checks as the socket path (gptserver.py:816-821)
            raise ValueError("Cannot generate less than 1 sample!")
        if self.n_nodes and n_samples < self.n_nodes:
            warnings.warn(f"Generating less samples ({n_samples}) than nodes ({self.n_nodes}) will not be efficient!")
        self.ring.pipe.set_sampling(sampling or self.sampling)  # sampling lives on the starter only
        idx = self.encode_prompts(prompt, n_samples)
        S = self.model.max_seq_length
        if any(max_new_tokens + p.numel() > S for p in idx):
            raise ValueError(f"Cannot generate {max_new_tokens} tokens - would exceed block size!")
        secondaries = self.node_config.get("nodes", {}).get("secondary", []) if self.n_nodes and self.n_nodes > 1 else []
        return RingSession(self.ring, secondaries, idx, max_new_tokens, mode=mode or self.decode_mode)

    def _starter_ring(self, n_samples: int, prompt: Any, max_new: int) -> Tuple[List[str], List[Tuple[int, float]]]:
        """Generation through the device ring.  Device-driven: everything is enqueued at once and the per-token
        timeline comes from device timestamps; host-fed: every sampled token is read back as it appears."""
        sess = self.open_ring_session(n_samples, prompt, max_new)
        tok_time: List[Tuple[int, float]] = [(0, 0.0)]
        try:
            if sess.mode == "host":
                t0 = [0.0]

                def on_token(slot: int, pos: int, tok: int) -> None:
                    if not t0[0]:
                        t0[0] = sess.t0_host
                    tok_time.append((len(tok_time), time.time() - t0[0]))
                    if self.on_token is not None:
                        self.on_token(slot, pos, tok)

                stats = sess.run(on_token=on_token)
            else:
                stats = sess.run()
                tok_time += [(i + 1, t) for i, t in enumerate(self.ring.token_times())]
            self.last_ring_stats = stats
            samples = sess.tokens()
        except BaseException:
            # Ctrl-C, a node that answered with an error, a tripped watchdog: whatever is still queued on the GPUs of the
            # ring must not spin out its own watchdog budget — poison every node's flags, then let the error surface
            try:
                sess.abort()
            except Exception:  # noqa: BLE001
                pass
            raise
        finally:
            sess.close()
        lens = {i: n for i, n in enumerate(sess.prompt_lens)}
        self.last_result = GenerationResult(samples=samples, prompt_lengths=lens, tok_time=tok_time,
                                            n_tokens=max_new * n_samples, elapsed=tok_time[-1][1])
        self.tok_time = tok_time
        truncated = [find_eot(samples[i], self.stop_tokens, lens[i]) for i in sorted(samples)]
        assert self.tok is not None
        return [self.tok.decode(s) for s in truncated], tok_time

    def _secondary_loop(self) -> None:
        assert self.runner is not None and self.transport is not None
        with catch_loop_errors(running_event=self.running):
            secondary_loop(self.runner, self.transport, self.running, n_samples=self.n_samples,
                           use_kv_cache=self.use_kv_cache)

    def stop_generation(self) -> int:
        try:
            self.running.clear()
            if self.ring is not None:
                if "starter" not in self.role:
                    self.ring.abort()  # a node told to stop while steps are queued: poison, so the ring drains
                self.ring.close()
                self.ring = None
            if "starter" not in self.role and self.inference_thread is not None \
                    and self.inference_thread is not threading.current_thread():
                self.inference_thread.join(timeout=10)
            if self.transport is not None:
                self.transport.shutdown()
                self.transport = None
            return 1
        except Exception:  # noqa: BLE001
            return 0

    def shutdown(self) -> int:
        try:
            ok = self.stop_generation()
            self.stop_webserv()
            return int(bool(ok))
        except Exception:  # noqa: BLE001
            return 0

    # ---- REST API (paths are tuples of segments, bodies raw bytes) ---------------------------------
    def GET(self, path: Tuple[str, ...], body: bytes) -> str:  # noqa: N802
        if len(path) == 0:
            return json.dumps(self.node_config)
        if path[0] == "status":
            return json.dumps({"role": self.role, "running": self.running.is_set(),
                               "model_loaded": self.model is not None, "device": self.model_device,
                               "error": repr(self.loop_error) if self.loop_error else None})
        raise HTTPError(404, "Not found")

    def POST(self, path: Tuple[str, ...], body: bytes) -> Optional[Dict[str, Any]]:  # noqa: N802
        if "secondary" not in (self.node_type or "secondary"):
            raise HTTPError(403, "Unable to initialize node!")
        if path and path[0] == "ring":
            if self.ring is None:
                raise HTTPError(409, "this node was not initialised for the device ring")
            try:
                return self.ring.handle(safe_loads(body))
            except Exception as e:  # noqa: BLE001
                self.loop_error = e
                raise HTTPError(500, f"ring op failed: {e!r}") from e
        if self._initializing.acquire(blocking=False) is False:
            raise HTTPError(409, "node initialisation already in progress")
        try:
            return self._post_init(path, body)
        finally:
            self._initializing.release()

    def _post_init(self, path: Tuple[str, ...], body: bytes) -> Optional[Dict[str, Any]]:
        if self.model is not None:
            raise HTTPError(403, f"Failed to configure node - the model was already initialized: {self.node_type}")
        if not path or path[0] != "init":
            raise HTTPError(404, "Not found")
        if self.running.is_set():
            raise HTTPError(409, "node is already running")
        init_msg = safe_loads(body)
        self.prev_node, self.next_node = init_msg["prev_node"], init_msg["next_node"]
        self.model_config = Config.from_dict(init_msg["model_config"])
        self.n_nodes = init_msg["n_nodes"]
        self.n_layers_local = init_msg["n_local_layers"]
        self.max_seq_length = init_msg.get("max_seq_length")
        self.n_samples = init_msg["n_samples"]
        if init_msg.get("sampling"):
            self.sampling = SamplingParams(**init_msg["sampling"])
        self.use_kv_cache = bool(init_msg.get("use_kv_cache", True))
        self.head_on = init_msg.get("head_on", "starter")
        transport = init_msg.get("transport", "socket")
        self.weights = init_msg.get("weights", self.weights)
        self.max_prompt_len = int(init_msg.get("max_prompt_len") or self.max_prompt_len or 0)
        if init_msg.get("watchdog_s"):
            self.watchdog_s = float(init_msg["watchdog_s"])
        if init_msg.get("random_init") is not None:
            self.random_init, self.stage_spec = int(init_msg["random_init"]), init_msg.get("stage_spec")
        elif init_msg.get("stage_spec") and self.stage_spec is None:
            self.stage_spec = init_msg["stage_spec"]
        if transport in ("p2p", "nccl"):
            from .ring import ring_capable

            if not ring_capable(self.model_device, self.ptdtype, self.model_config):
                raise HTTPError(400, f"node {self.role} ({self.model_device}, {self.dtype}) cannot join a {transport} device "
                                     "ring: restart the run with --transport socket")
        params = init_msg.pop("params", None)
        if params is not None:
            self._init_model(self.n_layers_local, model_parameters=params)
            del params
            gc.collect()
        elif self.random_init is not None:
            self._init_model(self.n_layers_local)
        else:
            if self.model_path is None:
                raise HTTPError(400, "The received message did not contain the model parameters - please "
                                     "specify a model chunk path when initializing GPTServer object")
            self._init_model(self.n_layers_local, model_path=self.model_path)
        logger_wp.info("Received initialization information!")
        if transport in ("p2p", "nccl"):
            rank = 1 + int(self.role.split(":")[1])
            handles = self.ring_setup(int(self.n_samples), rank, int(self.n_nodes), transport)
            return {"transport": transport, "handles": handles}
        self.inference_thread = threading.Thread(target=self.start_inference, daemon=True, args=(self.n_samples,))
        self.inference_thread.start()
        return {"transport": "socket"}

    def PUT(self, path: Tuple[str, ...], body: bytes) -> None:  # noqa: N802
        if self.node_type == "starter":
            raise HTTPError(501, "PUT not implemented!")
        if not path or path[0] != "stop":
            raise HTTPError(404, "Not found!")
        threading.Thread(target=self.shutdown, daemon=True).start()
        logger_wp.info("Received stopping directive")

    def DELETE(self, path: Tuple[str, ...], body: bytes) -> None:  # noqa: N802
        raise HTTPError(501, "DELETE not implemented!")
