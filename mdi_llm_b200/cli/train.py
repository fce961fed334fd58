#!/usr/bin/env python3
"""Trainer for litGPT-format models with data parallelism over NCCL.

Parity: reference ``src/train.py`` (:58-374; flags :377-477): memory-mapped ``train.bin`` /
``val.bin``, ``--init scratch|resume|hf``, GPT-NeoX style init (:35-55), tied embeddings
(:231-232), ``DistributedDataParallel`` under ``torchrun`` with gradient sync only on the last
micro-step (:88-103,250-251,325-328) — the one collective of the reference —, fp16 ``GradScaler``,
fused AdamW, cosine LR schedule, gradient accumulation + clipping, periodic evaluation,
checkpointing ``lit_model.pth`` + ``train_ckpt.pkl`` (:301-311), ``--patience`` early stop,
``--force-old`` on resume, MFU logging.

B200 notes: one process per GPU, NCCL over NVLink/NVSwitch (gloo on CPU-only hosts, used by the
tests); bf16 autocast by default on CUDA; MFU is reported against the *measured* sustained bf16
peak of ``MEASURED_PEAKS.json`` rather than the A100's 312 TF; the reference's ``get_num_params``
bug (``transformer.wpe`` on rope models, model.py:342-346) is not reproduced.

    torchrun --standalone --nproc-per-node 8 -m mdi_llm_b200.cli.train --ckpt checkpoints/custom/NanoLlama \
        --dataset data/shakespeare --init scratch --max-iters 1000
"""
from __future__ import annotations

import argparse
import math
import os
import pickle
import time
from contextlib import nullcontext
from pathlib import Path
from typing import Any, Dict


def build_parser() -> argparse.ArgumentParser:
    from .. import config as C

    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("-v", "--verb", action="store_true")
    p.add_argument("-c", "--compile", action="store_true", help="accepted for compatibility (not used)")
    p.add_argument("--ckpt", type=Path, default=Path("./checkpoints/custom/NanoLlama/"),
                   help="checkpoint directory: model_config.yaml (+ lit_model.pth / train_ckpt.pkl when resuming)")
    p.add_argument("--dataset", type=Path, default=Path("./data/shakespeare"), help="directory with train.bin and val.bin")
    p.add_argument("--init", type=str, default="scratch", choices=["scratch", "resume", "hf"])
    p.add_argument("--model-name", type=str, default=None, help="registry name used when --ckpt has no model_config.yaml")
    p.add_argument("-F", "--force-old", action="store_true", help="on resume keep the old training settings")
    p.add_argument("--batch-size", type=int, default=10)
    p.add_argument("--block-size", type=int, default=None, help="training context (default: model block size)")
    p.add_argument("--max-iters", type=int, default=100)
    p.add_argument("--patience", type=int, default=None)
    p.add_argument("--ckpt-interval", type=int, default=20)
    p.add_argument("-au", "--always-update", action="store_true")
    p.add_argument("--log-interval", type=int, default=10)
    p.add_argument("--eval-iters", type=int, default=20)
    p.add_argument("--grad-acc-steps", type=int, default=10)
    p.add_argument("--learning-rate", type=float, default=C.LEARNING_RATE)
    p.add_argument("--warmup-iters", type=int, default=None)
    p.add_argument("--device", type=str, default=C.DEVICE)
    p.add_argument("--dtype", type=str, default=None)
    p.add_argument("--hf-token", type=str, default=os.getenv("HF_TOKEN"))
    p.add_argument("--seed", type=int, default=10137)
    return p


def init_weights_neox(model, n_layer: int, n_embd: int) -> None:
    """GPT-NeoX initialisation: N(0, sqrt(2/(5 d))) everywhere, output projections scaled by
    depth (reference train.py:35-55)."""
    import torch.nn as nn

    std = math.sqrt(2.0 / 5 / n_embd)
    for name, mod in model.named_modules():
        if isinstance(mod, nn.Embedding):
            nn.init.normal_(mod.weight, mean=0.0, std=std)
        elif isinstance(mod, nn.Linear):
            out_proj = name.endswith("attn.proj") or name.endswith("mlp.proj")
            nn.init.normal_(mod.weight, mean=0.0, std=(1 / math.sqrt(n_embd) / n_layer) if out_proj else std)
            if mod.bias is not None:
                nn.init.zeros_(mod.bias)


def measured_peak_flops() -> float:
    import json

    for cand in (Path.cwd() / "MEASURED_PEAKS.json", Path(__file__).resolve().parents[2] / "MEASURED_PEAKS.json"):
        if cand.is_file():
            try:
                return float(json.loads(cand.read_text())["bf16_tflops_sustained"]) * 1e12
            except Exception:  # noqa: BLE001
                pass
    return 1.4e15  # profiling recipe's sustained fallback


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    import numpy as np
    import torch
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP

    from .. import config as C
    from ..models.config import Config
    from ..models.gpt import GPT
    from ..utils.checkpoint import load_sd
    from ..utils.data_loader import get_batch
    from ..utils.misc import estimate_loss, get_lr

    # ---- distributed setup ----------------------------------------------------------------------
    ddp = int(os.environ.get("RANK", -1)) != -1
    device = args.device
    if ddp:
        rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
        use_cuda = "cuda" in device and torch.cuda.is_available()
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=C.BACKEND if use_cuda else "gloo")
        if use_cuda:
            device = f"cuda:{local_rank}"
            torch.cuda.set_device(device)
        master = rank == 0
        if args.grad_acc_steps % world == 0:
            args.grad_acc_steps //= world  # same global batch as the single-process run
    else:
        rank, world, master = 0, 1, True
    torch.manual_seed(args.seed + rank)
    device_type = "cuda" if "cuda" in device else "cpu"
    dtype_name = args.dtype or (C.default_dtype() if device_type == "cuda" else "float32")
    ptdtype = C.DTYPE_TORCH_MAPPING[dtype_name]
    ctx = nullcontext() if device_type == "cpu" and ptdtype == torch.float32 else torch.autocast(device_type=device_type, dtype=ptdtype)

    # ---- data -------------------------------------------------------------------------------------
    train_data = np.memmap(args.dataset / "train.bin", dtype=np.uint16, mode="r")
    val_data = np.memmap(args.dataset / "val.bin", dtype=np.uint16, mode="r")

    # ---- model ------------------------------------------------------------------------------------
    ckpt_dir: Path = args.ckpt
    ckpt_dir.mkdir(parents=True, exist_ok=True)
    ckpt_model, ckpt_state = ckpt_dir / "lit_model.pth", ckpt_dir / "train_ckpt.pkl"
    iter_num, best_val_loss = 0, float("inf")
    settings: Dict[str, Any] = dict(learning_rate=args.learning_rate, max_iters=args.max_iters, batch_size=args.batch_size,
                                    grad_acc_steps=args.grad_acc_steps, warmup_iters=args.warmup_iters if args.warmup_iters is not None
                                    else max(1, args.max_iters // 20), lr_decay_iters=args.max_iters, min_lr=args.learning_rate / 10,
                                    weight_decay=C.WEIGHT_DECAY, beta1=C.BETA1, beta2=C.BETA2, grad_clip=C.GRAD_CLIP)
    opt_state = None
    if (ckpt_dir / "model_config.yaml").is_file():
        config = Config.from_file(ckpt_dir / "model_config.yaml")
    elif args.model_name:
        config = Config.from_name(args.model_name)
    else:
        raise FileNotFoundError(f"{ckpt_dir} has no model_config.yaml: pass --model-name")
    if args.block_size:
        config.block_size = args.block_size
    model = GPT(config)
    if args.init == "scratch":
        init_weights_neox(model, config.n_layer, config.n_embd)
    elif args.init == "resume":
        model.load_state_dict(load_sd(ckpt_model, "cpu"), strict=not config.tie_embeddings)
        with open(ckpt_state, "rb") as f:
            state = pickle.load(f)
        iter_num, best_val_loss, opt_state = state["iter_num"], state["best_val_loss"], state["optimizer"]
        if args.force_old:
            settings.update(state.get("train_settings", {}))
    else:  # "hf": weights converted beforehand (prepare_model.py) or downloaded now
        if not ckpt_model.is_file():
            from ..utils.download import download_from_hub

            download_from_hub(repo_id=f"{config.hf_config['org']}/{config.hf_config['name']}", access_token=args.hf_token,
                              checkpoint_dir=ckpt_dir.parent.parent)
        model.load_state_dict(load_sd(ckpt_model, "cpu"), strict=not config.tie_embeddings)
    if C.TrainingConfig.tie_embeddings and not config.tie_embeddings:
        model.transformer.wte.weight = model.lm_head.weight  # weight tying, every init mode (train.py:231-232)
    model.to(device)
    config.save(ckpt_dir)

    decay = [p for n, p in model.named_parameters() if p.dim() >= 2]
    no_decay = [p for n, p in model.named_parameters() if p.dim() < 2]
    optimizer = torch.optim.AdamW([{"params": decay, "weight_decay": settings["weight_decay"]},
                                   {"params": no_decay, "weight_decay": 0.0}], lr=settings["learning_rate"],
                                  betas=(settings["beta1"], settings["beta2"]), fused=device_type == "cuda")
    if opt_state is not None:
        optimizer.load_state_dict(opt_state)
    scaler = torch.amp.GradScaler(device_type, enabled=(ptdtype == torch.float16))
    raw_model = model
    if ddp:
        model = DDP(model, device_ids=[int(device.split(":")[1])] if device_type == "cuda" else None)

    peak = measured_peak_flops()
    tokens_per_iter = settings["grad_acc_steps"] * world * settings["batch_size"] * config.block_size
    if master:
        print(f"{raw_model.get_num_params() / 1e6:.2f} M parameters | {tokens_per_iter} tokens/iter | world {world} | {dtype_name} on {device}")

    # ---- loop -------------------------------------------------------------------------------------
    X, Y = get_batch(train_data, settings["batch_size"], device, config)
    t0 = time.time()
    local_iter, count_loss_incr, running_mfu = 0, 0, -1.0
    max_iters = settings["max_iters"]
    while iter_num <= max_iters:
        lr = get_lr(iter_num, settings["learning_rate"], settings["min_lr"], settings["warmup_iters"], settings["lr_decay_iters"])
        for g in optimizer.param_groups:
            g["lr"] = lr
        if iter_num % args.ckpt_interval == 0 and master:
            losses = estimate_loss(raw_model, train_data, val_data, settings["batch_size"], device, ctx=ctx, eval_iters=args.eval_iters)
            print(f"step {iter_num}: train loss {losses['train']:.4f}, val loss {losses['val']:.4f}")
            if losses["val"] < best_val_loss or args.always_update:
                if losses["val"] < best_val_loss:
                    count_loss_incr = 0
                best_val_loss = min(best_val_loss, losses["val"])
                if iter_num > 0:
                    with open(ckpt_state, "wb") as f:
                        pickle.dump({"optimizer": optimizer.state_dict(), "train_settings": settings, "iter_num": iter_num,
                                     "best_val_loss": best_val_loss, "config": config.asdict()}, f)
                    torch.save(raw_model.state_dict(), ckpt_model)
                    print(f"Saving state to {ckpt_state} and model to {ckpt_model}")
            else:
                count_loss_incr += 1
        if ddp and args.patience is not None:  # every rank must agree on stopping
            flag = torch.tensor([count_loss_incr if master else 0], device=device if device_type == "cuda" else "cpu")
            dist.broadcast(flag, 0)
            count_loss_incr = int(flag.item())
        if args.patience is not None and count_loss_incr >= args.patience:
            if master:
                print(f"No performance increase in the last {args.patience} evaluations - stopping!")
            break
        for micro in range(settings["grad_acc_steps"]):
            if ddp:  # all-reduce the gradients only once per optimizer step
                model.require_backward_grad_sync = micro == settings["grad_acc_steps"] - 1
            with ctx:
                logits = model(X)
                loss = torch.nn.functional.cross_entropy(logits.view(-1, logits.size(-1)).float(), Y.view(-1), ignore_index=-1)
                loss = loss / settings["grad_acc_steps"]
            X, Y = get_batch(train_data, settings["batch_size"], device, config)  # prefetch while the GPU works
            scaler.scale(loss).backward()
        if settings["grad_clip"]:
            scaler.unscale_(optimizer)
            torch.nn.utils.clip_grad_norm_(model.parameters(), settings["grad_clip"])
        scaler.step(optimizer)
        scaler.update()
        optimizer.zero_grad(set_to_none=True)
        dt, t0 = time.time() - t0, time.time()
        if iter_num % args.log_interval == 0 and master:
            lossf = loss.item() * settings["grad_acc_steps"]
            if local_iter >= 5:
                mfu = raw_model.estimate_mfu(settings["batch_size"] * settings["grad_acc_steps"] * world, dt, peak_flops=peak * world)
                running_mfu = mfu if running_mfu < 0 else 0.9 * running_mfu + 0.1 * mfu
            print(f"iter {iter_num}: loss {lossf:.4f}, lr {lr:.2e}, time {dt * 1000:.2f} ms, mfu {running_mfu * 100:.2f}%")
        iter_num += 1
        local_iter += 1
    if ddp:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
