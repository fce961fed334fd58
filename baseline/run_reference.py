"""Reference arm of ``bench.py``: the UNMODIFIED reference (``baseline/_ref``) through its own
public API (``sub.model_dist.GPTDistributed`` — what its ``starter.py`` / ``secondary.py`` call),
stock code path: loopback TCP sockets + pickle between nodes, CherryPy-style HTTP control, eager
PyTorch ops.  Nothing from ``mdi_llm_b200``'s models, kernels or engine is on that path; this
module only (a) prepares inputs in the reference's on-disk formats — a random-init litGPT
checkpoint of the benchmarked architecture, a synthetic ``tokenizer.json``, a loopback node JSON —
(b) provides import shims for three packages that are not installable offline (cherrypy,
accelerate, matplotlib; ``baseline/shims``), and (c) turns the reference's own per-token timeline
(``tok_time``, gptserver.py:902,952-956) into the benchmark metric.

One rank per GPU (torchrun): rank 0 is the starter, rank i the secondary i-1, each on
``cuda:LOCAL_RANK``.  Node counts the reference has no partition for (``N_LAYERS_NODES`` lacks the
entry, config.py:56-98 — e.g. 8 nodes) report ``unavailable`` with that reason.
"""
from __future__ import annotations

import json
import os
import sys
import time
import traceback
from pathlib import Path
from typing import Any, Dict, List

HERE = Path(__file__).resolve().parent
REF = HERE / "_ref"
SHIMS = HERE / "shims"
METRIC = "generated tokens/sec (whole box, device-timed, max over ranks) Llama-3-8B recurrent-pipeline decode"


def _unavailable(why: str, variant: str = "reference") -> Dict[str, Any]:
    return {"impl": variant, "unavailable": why}


def _write_tokenizer(ckpt: Path, vocab_size: int) -> None:
    """Synthetic HF ``tokenizer.json``: word-level vocabulary t0..t{V-1} + BOS/EOS specials."""
    from tokenizers import Tokenizer, models, pre_tokenizers

    vocab = {f"t{i}": i for i in range(vocab_size - 2)}
    vocab["<|begin_of_text|>"] = vocab_size - 2
    vocab["<|end_of_text|>"] = vocab_size - 1
    tk = Tokenizer(models.WordLevel(vocab, unk_token="t0"))
    tk.pre_tokenizer = pre_tokenizers.WhitespaceSplit()
    tk.save(str(ckpt / "tokenizer.json"))
    (ckpt / "tokenizer_config.json").write_text(json.dumps(
        {"bos_token": "<|begin_of_text|>", "eos_token": "<|end_of_text|>", "add_bos_token": False}))


def _write_checkpoint(ckpt: Path, model_name: str, device: str) -> None:
    """Random-init ``lit_model.pth`` + ``model_config.yaml`` in the litGPT layout (inputs only)."""
    import torch
    import yaml

    from sub.model import Config  # the reference's own Config (asdict -> model_config.yaml)

    cfg = Config.from_name(model_name)
    ckpt.mkdir(parents=True, exist_ok=True)
    g = torch.Generator(device=device).manual_seed(1234)

    def rnd(*shape: int) -> "torch.Tensor":
        return (torch.randn(*shape, generator=g, device=device, dtype=torch.float32) * 0.02).to(torch.bfloat16).cpu()

    def ones(n: int) -> "torch.Tensor":
        return (1.0 + 0.1 * torch.randn(n, generator=g, device=device)).to(torch.bfloat16).cpu()

    # parameter names and shapes come from the reference's own module tree (built on the meta device),
    # so every architecture its Config can describe gets a loadable checkpoint
    from sub.model import GPT

    with torch.device("meta"):
        skeleton = GPT(cfg)
    sd = {}
    for name, t in skeleton.state_dict().items():
        if t.ndim == 1 and name.endswith(".weight"):
            sd[name] = ones(t.shape[0])  # normalisation gains
        else:
            sd[name] = rnd(*t.shape)
    V = cfg.padded_vocab_size
    tmp = ckpt / "lit_model.pth.tmp"  # atomic: a run killed half-way must not leave a truncated checkpoint behind
    torch.save(sd, tmp)
    os.replace(tmp, ckpt / "lit_model.pth")
    with open(ckpt / "model_config.yaml", "w") as f:
        yaml.safe_dump(cfg.asdict(), f)
    _write_tokenizer(ckpt, V)


def _topology(n: int, base: int) -> Dict[str, Any]:
    def node(i: int) -> Dict[str, Any]:
        return {"addr": "127.0.0.1", "communication": {"port": base + i, "starter_addr": "127.0.0.1"},
                "inference": {"port_in": base + 8 + 2 * i, "port_out": base + 9 + 2 * i}, "device": f"cuda:{i}"}

    return {"nodes": {"starter": node(0), "secondary": [node(i) for i in range(1, n)]}}


def run_reference(args: Any, variant: str = "reference") -> Dict[str, Any]:
    """``variant``: ``"reference"`` = byte-stock reference (sockets + pickle, its own partition table);
    ``"reference-table"`` = the same plus, for node counts its table lacks, an injected ``N_LAYERS_NODES`` entry
    (data only, reference-style uniform split) so that e.g. 8 GPUs get a comparator; ``"reference-nccl"`` = the
    reference NCCL-p2p build: ``sub.connections`` replaced by ``baseline/nccl_connections.py`` (torch.distributed
    send/recv), everything else stock (table entry injected when missing)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not (REF / "sub" / "model_dist.py").is_file():
        return _unavailable("baseline/_ref missing: run `python baseline/install_reference.py` (needs /root/reference)") \
            if rank == 0 else {}
    for p in (str(SHIMS), str(REF)):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.chdir(REF)  # the reference resolves a few paths relative to its script directory
    try:
        import torch
        import torch.distributed as dist

        cpu_mode = os.environ.get("MDI_REF_DEVICE", "") == "cpu"  # plumbing test without a GPU
        if not cpu_mode:
            torch.cuda.set_device(local_rank)
        dev = "cpu" if cpu_mode else f"cuda:{local_rank}"
        dtype = "float32" if cpu_mode else "bfloat16"
        nccl = variant == "reference-nccl" and world > 1
        if world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            if nccl:  # the data plane of this variant; also used for the barriers around the run
                if cpu_mode:
                    dist.init_process_group("gloo")  # protocol test of the shim on a machine without GPUs
                else:
                    dist.init_process_group("nccl", device_id=torch.device(dev))
                import importlib.util

                spec = importlib.util.spec_from_file_location("sub.connections", HERE / "nccl_connections.py")
                mod = importlib.util.module_from_spec(spec)
                import sub  # noqa: F401  (the reference package: parent of the module being replaced)

                spec.loader.exec_module(mod)
                sys.modules["sub.connections"] = mod
                mod._edge_groups()  # collective: every rank creates the edge groups in the same order, now
            else:
                dist.init_process_group("gloo")  # only used here for barriers around the reference run
        from sub.config import N_LAYERS_NODES  # noqa: E402
        from sub.model import Config  # noqa: E402
        from sub.model_dist import GPTDistributed  # noqa: E402

        model_name = os.environ.get("MDI_REF_MODEL") or ("tiny-llama-1.1b" if getattr(args, "tiny", False) else args.model)
        cfg = Config.from_name(model_name)
        injected = None
        if variant != "reference" and world > 1 and (world not in N_LAYERS_NODES or cfg.n_layer not in N_LAYERS_NODES[world]):
            # data only: a partition entry in the reference's own format, uniform like its existing rows
            sec = -(-cfg.n_layer // world)  # ceil
            start = cfg.n_layer - sec * (world - 1)
            while start < 1:
                sec -= 1
                start = cfg.n_layer - sec * (world - 1)
            injected = {"N_LAYERS_START": start, "N_LAYERS_SECONDARY": sec}
            N_LAYERS_NODES.setdefault(world, {})[cfg.n_layer] = injected
        if world not in N_LAYERS_NODES or cfg.n_layer not in N_LAYERS_NODES[world]:
            return _unavailable(f"the reference has no layer partition for {world} nodes x {cfg.n_layer} layers "
                                f"(KeyError in N_LAYERS_NODES, src/sub/config.py:56-98)") if rank == 0 else {}
        n_samples = args.n_samples or world
        n_tokens = args.warmup + args.steps + 1
        seq_len = args.seq_len or ((args.prompt_len + n_tokens + 64) // 64) * 64
        ckpt = Path(os.environ.get("MDI_REF_CKPT_DIR", "/tmp/mdi_ref_ckpt")) / "custom" / (cfg.name + "-random")
        transport_desc = ("torch.distributed send/recv over NCCL replacing sub.connections (baseline/nccl_connections.py); "
                          "everything else stock") if nccl else "loopback TCP + pickle (reference stock path)"
        if rank == 0 and not (ckpt / "lit_model.pth").is_file():
            _write_checkpoint(ckpt, model_name, dev)
        chunk_dir = ckpt / "chunks" / f"{world}nodes"
        wanted = [chunk_dir / "model_starter.pth"] + [chunk_dir / f"model_secondary{i}.pth" for i in range(world - 1)]
        done_mark = chunk_dir / ".complete"
        if rank == 0 and world > 1 and not (done_mark.is_file() and all(f.is_file() for f in wanted)):
            import shutil

            shutil.rmtree(chunk_dir, ignore_errors=True)  # leftovers of an interrupted split
            # the reference's documented workflow: prepare_model.py splits the checkpoint first
            # (src/prepare_model.py:57-58).  (Its split-on-the-fly path calls torch.load(device=...),
            # model_dist.py:456, which current torch rejects.)
            from sub.utils import load_from_pt, split_and_store

            _, full_sd = load_from_pt(ckpt)
            split_and_store(full_sd, world, ckpt)
            del full_sd
            done_mark.write_text("ok")
        if world > 1:
            dist.barrier()
        # ports derived from the rendezvous port AND the node count, so that concurrent / stale runs never collide: the
        # reference binds its data sockets without SO_REUSEADDR (connections.py:122,292), and a scaling sweep runs
        # N = 1, 2, 4, 8 within TIME_WAIT of each other
        # (a block of 24 ports per (rendezvous port, node count): 8 control + 16 data)
        topo = _topology(world, 20000 + (int(os.environ.get("MASTER_PORT", "29500")) % 400) * 100
                         + {1: 0, 2: 24, 4: 48, 8: 72}.get(world, 0))
        topo_file = ckpt.parent / f"nodes_{world}.json"
        if rank == 0:
            topo_file.write_text(json.dumps(topo))
        if world > 1:
            dist.barrier()

        if rank > 0:
            node = GPTDistributed(f"secondary:{rank - 1}", topo_file, ckpt_dir=ckpt, device=dev, dtype=dtype, verb=False)
            node.start()  # blocks until the starter's PUT /stop
            if world > 1:
                dist.barrier()
            return {}

        prompt = " ".join(f"t{7 + 13 * i}" for i in range(args.prompt_len))
        t_setup = time.time()
        starter = GPTDistributed("starter", topo_file, ckpt_dir=ckpt, device=dev, dtype=dtype,
                                 model_seq_length=seq_len, verb=False, plots=True)
        setup_s = time.time() - t_setup
        tok_time: List[Any] = starter.start(n_samples=n_samples, tokens_per_sample=n_tokens, prompt=prompt)
        if world > 1:
            dist.barrier()
        # tok_time[i] = (i tokens generated in total, seconds since the loop started)
        lo, hi = args.warmup * n_samples, (args.warmup + args.steps) * n_samples
        if not tok_time or len(tok_time) <= hi:
            return _unavailable(f"reference produced {len(tok_time) if tok_time else 0} timeline points, need {hi + 1}")
        dt = tok_time[hi][1] - tok_time[lo][1]
        value = (hi - lo) / dt
        return {
            "metric": METRIC.replace("Llama-3-8B", cfg.name) if cfg.name != "Llama-3-8B" else METRIC, "value": round(value, 3), "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dt * 1e3 / args.steps, 5), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if dtype == "bfloat16" else dtype,
            "data": "synthetic prompts, random-init weights", "impl": variant,
            # `config`: identical keys / values to the ours arm (bench.py); how this arm runs it is under `details`
            "config": {"model": cfg.name, "global_batch": n_samples, "seq_len": seq_len, "prompt_len": args.prompt_len,
                       "parallelism": f"pp{world}", "tokens_per_step": n_samples,
                       "sampling": {"temperature": 0.8, "top_k": 200},
                       "l2_policy": "inputs (stage weights) larger than L2, no flush"},
            "details": {"partition": "reference table split" + (f", INJECTED table entry {injected}" if injected else ""),
                        "transport": transport_desc, "compute": "eager PyTorch / cuBLAS",
                        "sampling_source": "the reference's defaults (TOP_K = 200, TEMPERATURE = 0.8, config.py:47-52)",
                        "timing": "the reference's own per-token wall-clock timeline on the starter (tok_time)",
                        "shims": ["cherrypy", "accelerate", "matplotlib"], "setup_s": round(setup_s, 1)},
            "e2e": {"value": round(value, 3), "unit": "tokens/s", "note": "reference timing is end-to-end by construction"},
            "gpu_launches": 0,
        }
    except Exception as e:  # noqa: BLE001
        traceback.print_exc()
        return _unavailable(f"reference run failed: {type(e).__name__}: {e}"[:300]) if rank == 0 else {}
