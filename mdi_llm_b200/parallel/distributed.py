"""Orchestration API: ``GPTDistributed`` — what ``starter.py`` / ``secondary.py`` instantiate.

Parity: reference ``src/sub/model_dist.py`` ``GPTDistributed`` — constructor signature and
path resolution (:136-339: checkpoint dir, ``chunks/<N>nodes/`` lookup, on-the-fly split when
the chunks are absent, ``--chunk`` override, truncated ``model_seq_length``), ``start``
(:341-397), ``configure_nodes`` (:402-484: one ``POST /init`` per secondary with role,
prev/next node, model config, ``n_nodes``, ``n_local_layers``, ``n_samples``, ``max_seq_length``
and optionally the chunk itself), ``stop_nodes`` (:486-497), ``_request_to_node`` (:499-573).

Differences: ``n_local_layers`` is per secondary (non-uniform plans for topologies the reference
table lacks, e.g. 8 stages); ``start`` does not crash when plotting is off (the reference
indexes an empty ``time_gen``, model_dist.py:383, and then never stops its secondaries).

B200 data plane: where the reference's nodes open sockets at init (gptserver.py:540-583), ``configure_nodes``
here wires the *device ring* when the topology allows it (``transport="auto"|"p2p"|"nccl"``, see
:mod:`.server`): the starter allocates its hop buffers, every ``POST /init`` answers with the CUDA-IPC
handles of that secondary's buffers, and one ``POST /ring {"op": "connect"}`` per node hands it the handles
of its successor.  ``partition="half"`` plans half-layer units, ``weights="fp8"`` serves block-scaled fp8,
``random_init=<seed>`` replaces checkpoint chunks by synthetic weights (benchmarks).  ``open_session`` exposes
the prepared generation so that a caller can run it in timed segments (``bench.py``).
"""
from __future__ import annotations

import json
import warnings
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import torch

from ..models.config import Config
from ..models.partition import (chunk_dir, count_transformer_blocks, plan_layers, split_and_store,
                                stage_specs)
from ..utils.checkpoint import lazy_load, load_from_pt
from .control import call_node, request_to_node
from .server import GPTServer

FileType = Union[str, Path]
__all__ = ["GPTDistributed"]


class GPTDistributed:
    init_msg: Dict[str, Any] = {
        "role": "", "prev_node": {}, "next_node": {}, "model_config": {}, "n_nodes": 0,
        "n_samples": 0, "max_seq_length": None,
    }

    def __init__(
        self,
        node_type: str,
        config_file: Union[FileType, Dict[str, Any]],
        *,
        ckpt_dir: Optional[FileType] = None,
        chunk_path: Optional[FileType] = None,
        device: Optional[str] = None,
        dtype: Optional[str] = None,
        secondary_index: Optional[int] = None,
        model_seq_length: Optional[int] = None,
        **kwargs: Any,
    ) -> None:
        self.ckpt_dir = Path(ckpt_dir) if ckpt_dir is not None else None
        self.chunk_path = Path(chunk_path) if chunk_path is not None else None
        self.torch_device = device if device else None
        self.verb = bool(kwargs.get("verb", False))
        self.plots = bool(kwargs.get("plots", False))
        self.dtype = dtype
        self.partition_policy = kwargs.pop("partition", "auto")
        self.push_chunks = bool(kwargs.pop("push_chunks", False))
        self.head_on: str = kwargs.get("head_on", "starter")
        self.random_init: Optional[int] = kwargs.get("random_init")
        self.specs: Optional[List[Dict[str, Any]]] = None
        self.full_model_name = self.ckpt_dir.name if self.ckpt_dir else None
        self.node_type = node_type
        if isinstance(config_file, dict):
            self.node_config = config_file
        else:
            with open(config_file, "r") as f:
                self.node_config = json.load(f)
        self.model_config: Optional[Config] = None
        self.model_seq_length: Optional[int] = None
        self.plan: Optional[List[int]] = None

        if self.node_type == "starter":
            assert self.ckpt_dir, "No model was specified!"
            self.n_secondary = len(self.node_config["nodes"].get("secondary", []))
            self.n_nodes = 1 + self.n_secondary
            self.own_config = self.node_config["nodes"]["starter"]
            if self.chunk_path:
                node_chunks_dir = self.chunk_path.resolve().parent
                self.model_was_split = True
            else:
                node_chunks_dir = chunk_dir(self.ckpt_dir, self.n_nodes)
                self.model_was_split = node_chunks_dir.is_dir() and (node_chunks_dir / "model_starter.pth").is_file()
            if self.random_init is not None:  # synthetic weights: only model_config.yaml is read
                self.model_config, _ = load_from_pt(self.ckpt_dir, config_only=True)
                self.specs = stage_specs(self.n_nodes, self.model_config, self.partition_policy)
                kwargs["stage_spec"] = self.specs[0]
                self.model_was_split = True
            elif not self.model_was_split and self.n_nodes > 1:
                if self.verb:
                    print("Chunks not found! Splitting the model")
                self.model_config, full_model = load_from_pt(self.ckpt_dir)
                assert full_model is not None
                sub = self.partition_policy in ("half", "third") and not self.model_config.parallel_residual \
                    and self.head_on == "starter"
                specs = stage_specs(self.n_nodes, self.model_config, self.partition_policy) if sub else None
                if specs is not None and specs[0].get("unit") == "third":
                    self.plan = None
                    node_chunks_dir = split_and_store(full_model, self.n_nodes, self.ckpt_dir, specs=specs, verb=self.verb)
                elif specs is not None and specs[0].get("unit") == "half":
                    self.plan = None
                    node_chunks_dir = split_and_store(full_model, self.n_nodes, self.ckpt_dir, units=[sp["units"] for sp in specs],
                                                      verb=self.verb)
                else:
                    self.plan = plan_layers(self.n_nodes, self.model_config.n_layer, self.model_config,
                                            policy="balanced" if self.partition_policy in ("half", "third") else self.partition_policy)
                    node_chunks_dir = split_and_store(full_model, self.n_nodes, self.ckpt_dir, plan=self.plan,
                                                      config=self.model_config, verb=self.verb, head_on=self.head_on)
                self.model_was_split = not self.push_chunks
            else:
                self.model_config, _ = load_from_pt(self.ckpt_dir, config_only=True)
            if model_seq_length and model_seq_length > self.model_config.block_size:
                raise ValueError(
                    f"The truncated sequence length {model_seq_length} should be lower or equal than "
                    f"the model's max sequence length {self.model_config.block_size}")
            self.model_seq_length = model_seq_length
            if not self.chunk_path:
                self.chunk_path = (node_chunks_dir / "model_starter.pth") if self.n_nodes > 1 \
                    else self.ckpt_dir / "lit_model.pth"
            self.node_chunks_dir = node_chunks_dir
            self.gpt_serv = GPTServer(
                node_config=self.node_config, node_type=self.node_type, model_config=self.model_config,
                chunk_path=self.chunk_path, tokenizer_dir=self.ckpt_dir, model_device=self.torch_device,
                dtype=dtype, model_type=self.full_model_name, model_seq_length=self.model_seq_length, **kwargs)
        elif "secondary" in self.node_type:
            # (the reference insists on a chunk path or a checkpoint directory here and notes in a FIXME that a secondary
            #  could be model-agnostic, model_dist.py:281-286: with neither, this node simply waits for a POST /init that
            #  carries the model config and the chunk itself)
            split_type = self.node_type.split(":")
            self.secondary_index = secondary_index if len(split_type) < 2 else int(split_type[1])
            assert self.secondary_index is not None
            self.node_type = f"secondary:{self.secondary_index}"
            self.n_nodes = None
            if "nodes" in self.node_config:
                self.own_config = self.node_config["nodes"]["secondary"][self.secondary_index]
                self.n_nodes = 1 + len(self.node_config["nodes"]["secondary"])
            else:
                self.own_config = self.node_config
            if self.ckpt_dir and self.n_nodes and self.chunk_path is None:
                self.chunk_path = chunk_dir(self.ckpt_dir, self.n_nodes) / f"model_secondary{self.secondary_index}.pth"
            elif not self.chunk_path and not self.n_nodes:
                warnings.warn("Missing info about total n. of nodes, cannot select correct chunk")
            if self.ckpt_dir and (self.ckpt_dir / "model_config.yaml").is_file():
                self.model_config, _ = load_from_pt(self.ckpt_dir, config_only=True)
            self.gpt_serv = GPTServer(
                node_config=self.node_config, node_type=self.node_type, model_config=self.model_config,
                chunk_path=self.chunk_path, model_device=self.torch_device, dtype=dtype, **kwargs)
        else:
            raise ValueError(f"unknown node type {node_type!r}")
        self.torch_device = self.gpt_serv.model_device

    # ---------------------------------------------------------------------------------------------
    def start(self, *, n_samples: Optional[int] = None, tokens_per_sample: Optional[int] = None,
              prompt: Optional[Union[str, Sequence[torch.Tensor]]] = None,
              quiet: bool = False) -> Optional[List[Tuple[int, float]]]:
        if self.node_type != "starter":
            try:
                self.gpt_serv.block()
            except KeyboardInterrupt:
                self.gpt_serv.shutdown()
                print("Node was stopped!")
            return None
        assert n_samples and tokens_per_sample and self.model_config
        try:
            if not self.configure_nodes(n_samples=n_samples):
                raise RuntimeError("Unable to initialize network nodes!")
            out_text, time_gen = self.gpt_serv.launch_starter(n_samples, tokens_per_sample, prompt)
            self.out_text = out_text
            if not quiet:
                print("-------------------------------------------------")
                print("Produced output:\n")
                for i, smpl in enumerate(out_text):
                    print("-------------------------------------------------")
                    print(f"Sample {i + 1}:")
                    print(smpl, "\n")
                print("-------------------------------------------------")
                if time_gen:
                    print(f"Total generation time: {time_gen[-1][1]}")
            return time_gen
        except KeyboardInterrupt:
            self.gpt_serv.shutdown()
            print("Node was stopped!")
            return None
        finally:
            self.stop_nodes()

    # ---------------------------------------------------------------------------------------------
    def _secondary_layer_counts(self) -> List[int]:
        """Layers of every secondary: from the chunk files if they are on this file system, else
        from the plan."""
        assert self.model_config is not None
        counts: List[int] = []
        for i in range(self.n_secondary):
            f = self.node_chunks_dir / f"model_secondary{i}.pth"
            if f.is_file():
                counts.append(count_transformer_blocks(lazy_load(f)))
            else:
                counts = []
                break
        if len(counts) == self.n_secondary:
            return counts
        if self.specs is not None:
            return [sp["n_blocks"] for sp in self.specs[1:]]
        plan = self.plan or plan_layers(self.n_nodes, self.model_config.n_layer, self.model_config,
                                        policy="balanced" if self.partition_policy in ("half", "third") else self.partition_policy)
        return list(plan[1:])

    def configure_nodes(self, n_samples: int) -> int:
        if self.node_type != "starter":
            raise ValueError("This method can only be called on starter nodes!")
        if not self.model_config:
            raise ValueError("The model configuration was not loaded!")
        nodes = self.node_config["nodes"]
        secondaries = nodes.get("secondary", [])
        serv = self.gpt_serv
        transport = serv.resolve_transport()
        self.transport = transport
        handles: List[Dict[str, Any]] = []
        if transport != "socket":  # device ring: the starter's hop buffers first (the last node stores into them)
            self.warn_if_plan_does_not_fit(n_samples)
            handles.append(serv.ring_setup(n_samples, 0, self.n_nodes, transport))
        if not secondaries:
            if self.verb:
                print("No secondary nodes found! Running standalone")
            if transport != "socket":
                serv.ring.connect(None)
            return 1
        counts = self._secondary_layer_counts()
        ring = [nodes["starter"]] + list(secondaries)  # ring order == order in the JSON
        s = serv.sampling
        for i, sec in enumerate(secondaries):
            msg = dict(self.init_msg)
            msg.update(
                role=f"secondary:{i}", model_config=self.model_config.asdict(), n_nodes=self.n_nodes,
                n_local_layers=counts[i], n_samples=n_samples, prev_node=ring[i],
                next_node=ring[(i + 2) % len(ring)], max_seq_length=self.model_seq_length,
                sampling=dict(temperature=s.temperature, top_k=s.top_k, top_p=s.top_p, seed=s.seed),
                use_kv_cache=serv.use_kv_cache, head_on=self.head_on, transport=transport, weights=serv.weights,
                max_prompt_len=serv.max_prompt_len, watchdog_s=serv.watchdog_s,
            )
            if self.random_init is not None:
                assert self.specs is not None
                msg.update(random_init=self.random_init, stage_spec=self.specs[i + 1])
            elif not self.model_was_split:
                msg["params"] = torch.load(self.node_chunks_dir / f"model_secondary{i}.pth",
                                           map_location="cpu", weights_only=True)
            addr = f"http://{sec['addr']}:{sec['communication']['port']}/init"
            status, body = call_node("post", addr, msg, verb=self.verb, timeout=3600.0)
            chunk_file = self.node_chunks_dir / f"model_secondary{i}.pth"
            if status == 400 and "params" not in msg and "model parameters" in str(body) and chunk_file.is_file():
                # a model-agnostic secondary (started without --ckpt / --chunk, or on a host that does not see this file
                # system): ship its chunk inside the init message, as the reference does after splitting on the fly
                # (model_dist.py:454-463)
                if self.verb:
                    print(f"Node secondary:{i} has no chunk of its own: sending {chunk_file.name} with the init message")
                msg["params"] = torch.load(chunk_file, map_location="cpu", weights_only=True)
                status, body = call_node("post", addr, msg, verb=self.verb, timeout=3600.0)
            if status != 200:
                print(f"Node secondary:{i} refused initialisation ({status}): {body}")
                return 0
            if transport != "socket":
                if not isinstance(body, dict) or "handles" not in body:
                    print(f"Node secondary:{i} did not return hop-buffer handles: {body!r}")
                    return 0
                handles.append(body["handles"])
        if transport != "socket":  # every node maps the buffers of its successor (ring order)
            for j, sec in enumerate(secondaries):
                addr = f"http://{sec['addr']}:{sec['communication']['port']}/ring"
                status, body = call_node("post", addr, {"op": "connect", "next": handles[(j + 2) % len(handles)]},
                                         max_n_requests=3, retry_wait=0.5, timeout=600.0)
                if status != 200:
                    print(f"Node secondary:{j} could not map its successor's buffers ({status}): {body}")
                    return 0
            serv.ring.connect(handles[1])
        return 1

    def warn_if_plan_does_not_fit(self, n_samples: int, capacity: Optional[int] = None) -> List[str]:
        """HBM budget of every stage of the plan (``models/memory.py``: weights + KV slots + hop buffers) against the
        device's memory, BEFORE any node allocates: a stage that cannot fit is announced with its numbers instead of an
        out-of-memory error half-way through a 100 GB chunk.  Advisory (a warning per stage), never fatal."""
        try:
            from ..models.memory import check_plan
            from ..models.partition import stage_specs

            assert self.model_config is not None
            if capacity is None:
                capacity = int(torch.cuda.get_device_properties(self.torch_device).total_memory)
            specs = self.specs or stage_specs(self.n_nodes, self.model_config, self.partition_policy)
            serv = self.gpt_serv
            msgs = check_plan(self.model_config, specs, n_samples, int(self.model_seq_length or self.model_config.block_size),
                              getattr(serv, "max_prompt_len", None) or None, getattr(serv, "weights", "bf16"), capacity=capacity)
        except Exception:  # noqa: BLE001  (an estimate must never stop a run)
            return []
        for m in msgs:
            warnings.warn(m)
        return msgs

    def open_session(self, n_samples: int, tokens_per_sample: int,
                     prompt: Optional[Union[str, Sequence[torch.Tensor]]] = None, mode: Optional[str] = None) -> Any:
        """Configure the nodes and return the prepared generation (:class:`~.ring.RingSession`): ``run(rounds)``
        enqueues prefill + that many decode rounds on every node and returns per-node device times, ``tokens()``
        the result.  Only for the device ring (``transport`` p2p / nccl); call :meth:`stop_nodes` when done."""
        if self.node_type != "starter":
            raise ValueError("This method can only be called on starter nodes!")
        if not self.configure_nodes(n_samples=n_samples):
            raise RuntimeError("Unable to initialize network nodes!")
        if self.gpt_serv.ring is None:
            raise RuntimeError(f"open_session needs the device ring; this topology resolved to transport {self.transport!r}")
        return self.gpt_serv.open_ring_session(n_samples, prompt, tokens_per_sample, mode=mode)

    def stop_nodes(self) -> int:
        out = 1
        for sec in self.node_config["nodes"].get("secondary", []):
            addr = f"http://{sec['addr']}:{sec['communication']['port']}/stop"
            out *= self._request_to_node("put", addr, "", max_n_requests=3)
        return out

    def _request_to_node(self, req_type: str, addr: str, content: Any, max_n_requests: int = 100) -> int:
        return request_to_node(req_type, addr, content, max_n_requests=max_n_requests, verb=self.verb)
