"""Legacy-generation tooling: checkpoint inspection / surgery / split check, memory
plots, tokenizer demo, offline splitter, toy models, split_map.json."""
import json
from pathlib import Path

import pytest
import torch

from mdi_llm_b200.cli import inspect_checkpoint, plot_mem, split_model, test_tokenizer as tok_cli
from mdi_llm_b200.models.bigram import BigramLanguageModel, TinyAttentionLM
from mdi_llm_b200.models.gpt import build_mask_cache
from mdi_llm_b200.models.partition import N_LAYERS_NODES
from mdi_llm_b200.utils.checkpoint import write_random_checkpoint


def test_split_map_json_mirrors_table():
    p = Path(inspect_checkpoint.__file__).resolve().parents[1] / "models" / "split_map.json"
    data = json.loads(p.read_text())
    assert {int(n): {int(l): v for l, v in t.items()} for n, t in data.items()} == N_LAYERS_NODES


def test_mask_cache():
    m = build_mask_cache(5)
    assert m.shape == (1, 1, 5, 5) and m.dtype == torch.bool
    assert m[0, 0, 2].tolist() == [True, True, True, False, False]


def test_inspect_lit_dir_and_split_roundtrip(tmp_path, tiny_llama_cfg, capsys):
    ck = write_random_checkpoint(tmp_path / "tiny", tiny_llama_cfg, dtype=torch.float32)
    keys = tmp_path / "keys.txt"
    assert inspect_checkpoint.main([str(ck), "--split", "2", "--save-keys", str(keys)]) == 0
    out = capsys.readouterr().out
    assert "kind: litgpt" in out and "wire round-trip ok" in out and f"transformer blocks: {tiny_llama_cfg.n_layer}" in out
    assert "transformer.wte.weight" in keys.read_text()


def test_model_surgery_fixes_dataset_name(tmp_path, capsys):
    d = tmp_path / "shakespeare_bpe" / "out"
    d.mkdir(parents=True)
    ck = d / "ckpt.pt"
    torch.save({"model": {"w": torch.zeros(2)}, "model_args": {"n_layer": 1}, "config": {"DATASET": "wrong"}}, ck)
    assert inspect_checkpoint.main([str(ck), "--fix-dataset"]) == 0
    assert "fixed dataset name" in capsys.readouterr().out
    assert torch.load(ck, weights_only=False)["config"]["DATASET"] == "shakespeare_bpe"
    assert inspect_checkpoint.main([str(ck), "--fix-dataset"]) == 0
    assert "dataset name ok" in capsys.readouterr().out


def test_plot_mem_collects_node_curves(tmp_path, capsys):
    d = tmp_path / "mem-usage" / "gpt2"
    d.mkdir(parents=True)
    for name, base in [("mem_2nodes_starter.csv", 100), ("mem_2nodes_secondary0.csv", 200), ("mem_3nodes_starter.csv", 1)]:
        (d / name).write_text("time_s,rss_mib,gpu0_used_mib\n" + "".join(f"{i * 0.5},{base + i},{i}\n" for i in range(4)))
    curves = plot_mem.collect(d, 2, "rss_mib")
    assert set(curves) == {"First node", "Node 2"} and curves["Node 2"][1][-1] == 203
    assert plot_mem.main(["gpt2", "2", "--logs", str(tmp_path / "mem-usage"), "-o", str(tmp_path / "m.png")]) == 0
    assert "peak 203" in capsys.readouterr().out


def test_tokenizer_demo_roundtrip(tmp_path):
    txt = tmp_path / "t.txt"
    txt.write_text("O, that this too too solid flesh would melt, thaw and resolve itself into a dew! " * 20)
    assert tok_cli.main(["--text", str(txt), "--vocab-size", "300", "--save", str(tmp_path / "tok")]) == 0
    assert tok_cli.main(["--text", str(txt), "--kind", "char", "--sentence", "solid flesh"]) == 0


def test_split_model_cli(tmp_path, tiny_llama_cfg):
    ck = write_random_checkpoint(tmp_path / "tiny", tiny_llama_cfg, dtype=torch.float32)
    assert split_model.main([str(ck), "--n-nodes", "2", "--head-on", "finisher"]) == 0
    st = torch.load(ck / "chunks/2nodes/model_starter.pth", weights_only=True)
    fin = torch.load(ck / "chunks/2nodes/model_secondary0.pth", weights_only=True)
    assert "lm_head.weight" not in st and "lm_head.weight" in fin and "transformer.ln_f.weight" in fin


def test_toy_models_train_a_little():
    torch.manual_seed(0)
    data = torch.randint(0, 11, (4, 17))
    for m in (BigramLanguageModel(11), TinyAttentionLM(11, n_embd=16, n_head=2, block_size=16)):
        opt = torch.optim.AdamW(m.parameters(), lr=1e-2)
        first = None
        for _ in range(30):
            _, loss = m(data[:, :-1], data[:, 1:])
            first = first if first is not None else loss.item()
            opt.zero_grad(); loss.backward(); opt.step()
        assert loss.item() < first
        assert m.generate(data[:1, :3], 5, temperature=0.0).shape == (1, 8)


def test_shared_parse_args_builds_both_flag_sets():
    from mdi_llm_b200.cli.common import parse_args

    tr = parse_args(True, ["--batch-size", "4", "--init", "resume", "--always-update"])
    assert tr.batch_size == 4 and tr.init == "resume" and tr.always_update and not tr.verb
    gen = parse_args(False, ["--n-samples", "3", "-p", "--n-tokens", "10", "--prompt", "hi"])
    assert gen.plots and gen.n_samples == 3 and gen.n_tokens == 10 and gen.prompt == "hi"


def test_tools_accept_the_reference_spellings(tmp_path, capsys):
    """mem_monitor (-o/--output, --img, one quoted command), plot_tok_time (positional MODEL_DIR, -nt) and prepare_owt
    (positional TOKENIZER_PATH, --data-path, --nproc) take the argument forms of the reference's scripts."""
    import sys

    from mdi_llm_b200.cli import mem_monitor, plot_tok_time, prepare_owt
    from mdi_llm_b200.cli.common import tokens_time_csv_name
    from mdi_llm_b200.text.tokenizer import write_bytes_tokenizer

    out, img = tmp_path / "mem" / "usage.csv", tmp_path / "mem" / "usage_plot.png"
    rc = mem_monitor.main(["-i", "0.05", "--output", str(out), "--img", str(img), f"{sys.executable} -c \"import time; time.sleep(0.3)\""])
    assert rc == 0
    rows = out.read_text().strip().splitlines()
    assert rows[0].startswith("time_s,rss_mib") and len(rows) >= 2

    logs = tmp_path / "logs"
    logs.mkdir()
    for k in (1, 2):
        (logs / tokens_time_csv_name(k, "NanoLlama", 3)).write_text("".join(f"{0.1 * i / k},{i}\n" for i in range(1, 6)))
    assert plot_tok_time.main([str(tmp_path / "ckpt" / "NanoLlama"), "-nt", "--n-samples", "3", "--logs-dir", str(logs)]) == 0
    assert "speed-up" in capsys.readouterr().out

    tok_dir, text_dir, data = tmp_path / "tok", tmp_path / "text", tmp_path / "owt"
    tok_dir.mkdir(); text_dir.mkdir()
    write_bytes_tokenizer(tok_dir)
    for i in range(3):
        (text_dir / f"d{i}.txt").write_text(f"document number {i} " * 20)
    assert prepare_owt.main([str(tok_dir), "--data-path", str(data), "--nproc", "1", "--text-dir", str(text_dir)]) == 0
    assert (data / "train.bin").is_file() and (data / "val.bin").is_file()


def test_inspect_lit_reference_form(tmp_path, tiny_llama_cfg, capsys, monkeypatch):
    """`inspect_lit --model <dir> -s` (scripts/inspect_lit.py:106-131): config dump, block-count check, key file."""
    from mdi_llm_b200.cli import inspect_lit

    ck = write_random_checkpoint(tmp_path / "tiny", tiny_llama_cfg, dtype=torch.float32)
    monkeypatch.chdir(tmp_path)
    assert inspect_lit.main(["--model", str(ck), "-s", "--device", "cpu"]) == 0
    out = capsys.readouterr().out
    assert f"{tiny_llama_cfg.n_layer} transformer blocks" in out
    assert "transformer.wte.weight" in (tmp_path / "tmp" / "tiny_params_keys_lit.txt").read_text()
    with pytest.raises(SystemExit):
        inspect_lit.main([])
    # scripts/test_tok.py: the special tokens of a checkpoint's tokenizer
    from mdi_llm_b200.text.tokenizer import write_bytes_tokenizer

    write_bytes_tokenizer(ck)
    assert tok_cli.main([str(ck)]) == 0
    assert "Beginning of sentence: 256" in capsys.readouterr().out


def test_prepare_data_legacy_forms(tmp_path):
    """`prepare_data.py DATA_DIR -t character` / `-t bpe --vocab-size N` of the older generations (old/GPT2/prepare_data.py:22-47)."""
    import numpy as np

    from mdi_llm_b200.cli import prepare_data

    d = tmp_path / "shakespeare"
    d.mkdir()
    (d / "input.txt").write_text("To be, or not to be, that is the question. " * 40)
    assert prepare_data.main([str(d), "-t", "character"]) == 0
    n_char = np.fromfile(d / "train.bin", dtype=np.uint16).size + np.fromfile(d / "val.bin", dtype=np.uint16).size
    assert n_char == len("To be, or not to be, that is the question. " * 40)
    out = tmp_path / "bpe"
    assert prepare_data.main([str(d), "-t", "bpe", "--vocab-size", "270", "--out-dir", str(out)]) == 0
    n_bpe = np.fromfile(out / "train.bin", dtype=np.uint16).size + np.fromfile(out / "val.bin", dtype=np.uint16).size
    assert 0 < n_bpe < n_char  # merges shorten the sequence
