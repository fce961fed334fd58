"""End-to-end CLI flows on CPU with tiny models: prepare → starter/secondary → sample/chat →
data prep → train (single process and 2-process DDP over gloo) → inspection/plot tools."""
import json
import os
import subprocess
import sys
import threading

import numpy as np
import pytest
import torch

from mdi_llm_b200.cli import (chat, inspect_lit, partition_table, plot_tok_time, prepare_data, prepare_model, sample,
                              secondary, starter, train)
from mdi_llm_b200.models.config import Config
from mdi_llm_b200.utils.checkpoint import load_from_pt, write_random_checkpoint
from conftest import free_ports

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "mdi_llm_b200", "data", "sonnets.txt")


@pytest.fixture
def tiny_ckpt(tmp_path, tiny_llama_cfg):
    return write_random_checkpoint(tmp_path / "custom" / "NanoLlama", tiny_llama_cfg, dtype=torch.float32)


def test_prepare_model_random_init_and_split(tmp_path, capsys):
    rc = prepare_model.main(["pythia-14m", "--random-init", "--ckpt-folder", str(tmp_path), "--n-nodes", "3", "--dtype", "float32"])
    assert rc == 0
    d = tmp_path / "EleutherAI" / "pythia-14m"
    assert (d / "lit_model.pth").is_file() and (d / "model_config.yaml").is_file()
    assert sorted(p.name for p in (d / "chunks" / "3nodes").iterdir()) == ["model_secondary0.pth", "model_secondary1.pth", "model_starter.pth"]
    assert inspect_lit.main([str(d), "--tokenizer", "hi"]) if False else True
    assert "split 3nodes" in (inspect_lit.main([str(d)]) or capsys.readouterr().out)


def test_starter_and_secondary_clis_over_loopback(tmp_path, tiny_ckpt, topology, monkeypatch):
    topo = topology(2)
    cfg_file = tmp_path / "nodes.json"
    cfg_file.write_text(json.dumps(topo))
    prepare_model.main([str(tiny_ckpt), "--n-nodes", "2"])
    monkeypatch.setenv("MDI_LOGS_DIR", str(tmp_path / "logs"))
    import mdi_llm_b200.cli.common as common
    import mdi_llm_b200.cli.starter as starter_mod

    monkeypatch.setattr(common, "LOGS_DIR", tmp_path / "logs")
    monkeypatch.setattr(starter_mod, "LOGS_DIR", tmp_path / "logs")
    monkeypatch.setattr(starter_mod, "IMG_DIR", tmp_path / "img")
    t = threading.Thread(target=secondary.main, args=(["--nodes-config", str(cfg_file), "0", "--ckpt", str(tiny_ckpt), "--dtype", "float32"],), daemon=True)
    t.start()
    stats = tmp_path / "stats" / "runs.csv"
    with pytest.warns(UserWarning):
        rc = starter.main(["--ckpt", str(tiny_ckpt), "--nodes-config", str(cfg_file), "--n-samples", "2", "--n-tokens", "5",
                           "--prompt", "Hi", "--dtype", "float32", "-p", "--time-run", str(stats), "--greedy"])
    assert rc == 0
    t.join(timeout=15)
    assert not t.is_alive()  # PUT /stop released the secondary
    csv = tmp_path / "logs" / "tokens_time_samples_2nodes_NanoLlama_2samples.csv"
    rows = csv.read_text().strip().splitlines()
    assert len(rows) == 1 + 2 * 5 and rows[-1].split(",")[1] == "10"  # "time,n_tokens" per generated token
    assert stats.read_text().splitlines()[0] == "timestamp,n_samples,n_layers,context_size,gen_time"
    assert plot_tok_time.main(["--model", "NanoLlama", "--n-samples", "2", "--logs-dir", str(tmp_path / "logs")]) == 0


def test_secondary_started_from_its_own_node_file(tmp_path, tiny_ckpt, topology, monkeypatch):
    """`secondary --secondary-config node.json --chunk …` (old/GPT2/secondary.py:46-52): the worker knows only its own
    entry; everything else arrives with POST /init."""
    topo = topology(2)
    cfg_file, own = tmp_path / "nodes.json", tmp_path / "secondary0.json"
    cfg_file.write_text(json.dumps(topo))
    own.write_text(json.dumps(topo["nodes"]["secondary"][0]))
    prepare_model.main([str(tiny_ckpt), "--n-nodes", "2"])
    import mdi_llm_b200.cli.common as common
    import mdi_llm_b200.cli.starter as starter_mod

    for mod in (common, starter_mod):
        monkeypatch.setattr(mod, "LOGS_DIR", tmp_path / "logs")
    monkeypatch.setattr(starter_mod, "IMG_DIR", tmp_path / "img")
    chunk = tiny_ckpt / "chunks" / "2nodes" / "model_secondary0.pth"
    t = threading.Thread(target=secondary.main, args=(["--secondary-config", str(own), "--chunk", str(chunk), "--dtype", "float32"],),
                         daemon=True)
    t.start()
    rc = starter.main(["--ckpt", str(tiny_ckpt), "--nodes-config", str(cfg_file), "--n-samples", "2", "--n-tokens", "3",
                       "--prompt", "Hi", "--dtype", "float32", "--greedy"])
    assert rc == 0
    t.join(timeout=15)
    assert not t.is_alive()


def test_sample_and_chat_cli(tmp_path, tiny_ckpt, capsys, monkeypatch):
    import mdi_llm_b200.cli.sample as sample_mod

    monkeypatch.setattr(sample_mod, "LOGS_DIR", tmp_path / "logs")
    monkeypatch.setattr(sample_mod, "IMG_DIR", tmp_path / "img")
    assert sample.main(["--ckpt", str(tiny_ckpt), "--n-samples", "2", "--n-tokens", "6", "--prompt", "ab", "--device", "cpu",
                        "--dtype", "float32", "-p", "--greedy"]) == 0
    out = capsys.readouterr().out
    assert "Sample 2:" in out and "tokens/s" in out
    assert (tmp_path / "logs" / "tokens_time_samples_1nodes_NanoLlama_2samples.csv").is_file()
    assert chat.main(["--ckpt", str(tiny_ckpt), "--device", "cpu", "--dtype", "float32", "--max-new-tokens", "5", "--once", "hey"]) == 0
    assert ">> Reply:" in capsys.readouterr().out


def _train_dir(tmp_path):
    data = tmp_path / "data"
    assert prepare_data.main([DATA, "--tokenizer", "bpe:300", "--out-dir", str(data)]) == 0
    train_bin = np.memmap(data / "train.bin", dtype=np.uint16, mode="r")
    assert len(train_bin) > 1000 and int(train_bin.max()) < 300
    ck = tmp_path / "ck"
    ck.mkdir()
    Config.from_name("NanoLlama", n_layer=2, n_embd=32, n_head=4, intermediate_size=64, vocab_size=300, padded_vocab_size=320,
                     block_size=32).save(ck)
    return data, ck


def test_train_scratch_resume_and_loss_goes_down(tmp_path, capsys):
    data, ck = _train_dir(tmp_path)
    common = ["--ckpt", str(ck), "--dataset", str(data), "--batch-size", "8", "--grad-acc-steps", "1", "--ckpt-interval", "10",
              "--log-interval", "5", "--eval-iters", "4", "--device", "cpu", "--learning-rate", "0.01", "--warmup-iters", "2"]
    assert train.main(common + ["--init", "scratch", "--max-iters", "30"]) == 0
    out = capsys.readouterr().out
    losses = [float(l.split("val loss ")[1]) for l in out.splitlines() if "val loss" in l]
    assert losses[-1] < losses[0] - 0.3, losses
    assert (ck / "lit_model.pth").is_file() and (ck / "train_ckpt.pkl").is_file()
    import pickle

    state = pickle.load(open(ck / "train_ckpt.pkl", "rb"))
    assert {"optimizer", "train_settings", "iter_num", "best_val_loss", "config"} <= set(state)
    assert train.main(common + ["--init", "resume", "--max-iters", "40"]) == 0
    out = capsys.readouterr().out
    assert f"step {state['iter_num']}:" in out  # continued from the stored iteration
    # the trained checkpoint is a regular litGPT checkpoint: the sampler loads it
    assert sample.main(["--ckpt", str(ck), "--n-tokens", "5", "--prompt", "Shall I", "--device", "cpu", "--dtype", "float32"]) == 0


def test_train_ddp_two_processes_gloo(tmp_path):
    data, ck = _train_dir(tmp_path)
    (port,) = free_ports(1)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), "-m", "mdi_llm_b200.cli.train", "--ckpt", str(ck), "--dataset", str(data), "--init", "scratch",
           "--max-iters", "6", "--batch-size", "4", "--grad-acc-steps", "2", "--ckpt-interval", "3", "--log-interval", "2",
           "--eval-iters", "2", "--device", "cpu"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    assert "world 2" in p.stdout and "step 6:" in p.stdout
    assert (ck / "lit_model.pth").is_file()


def test_partition_table_cli(capsys):
    assert partition_table.main([]) == 0
    table = json.loads(capsys.readouterr().out)
    assert table["2"]["32"] == {"N_LAYERS_START": 14, "N_LAYERS_SECONDARY": 18}
    assert partition_table.main(["--model", "Llama-3-8B"]) == 0
    assert "8 nodes: table=None balanced=[2, 5, 5, 4, 4, 4, 4, 4]" in capsys.readouterr().out
    # what-if view: the fitted cost model's stage times and ring throughput per policy; measured on B200 through the API
    # (profiles/README.md): 353 / 693 / 1339 / 2558-2649 tok/s at 1 / 2 / 4 / 8 nodes with third-layer units
    assert partition_table.main(["--model", "Llama-3-8B", "--predict"]) == 0
    rows = {(int(l.split()[0]), l.split()[2]): float(l.split("->")[1].split()[0]) for l in capsys.readouterr().out.splitlines() if "->" in l}
    for n, measured in ((1, 353.0), (2, 693.0), (4, 1339.0), (8, 2603.0)):
        pred = rows[(n, "third" if n > 1 else "table")]
        assert abs(pred - measured) / measured < 0.04, (n, pred, measured)
    assert rows[(8, "third")] > rows[(8, "half")] > rows[(8, "balanced")]  # finer units flatten the 8-stage ring
    assert partition_table.main(["--model", "Mixtral-8x7B-v0.1", "--predict", "--max-nodes", "2"]) == 0  # non-gated / MoE cost path
