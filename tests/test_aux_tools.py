"""Auxiliary tools that have no other coverage: memory monitor, download CLI (offline behaviour),
OpenWebText shard writer, optional-import helpers."""
import csv
import sys

import numpy as np
import pytest

from mdi_llm_b200.cli import download_weights, mem_monitor, prepare_owt
from mdi_llm_b200.utils.imports import LazyModule, ModuleAvailableCache, RequirementCache, module_available, requires


def test_mem_monitor_samples_a_command(tmp_path):
    out = tmp_path / "mem.csv"
    rc = mem_monitor.main(["-i", "0.05", "-o", str(out), "--", sys.executable, "-c", "import time; x = bytearray(30 << 20); time.sleep(0.4)"])
    assert rc == 0
    rows = list(csv.DictReader(open(out)))
    assert len(rows) >= 2 and float(rows[-1]["time_s"]) > 0
    assert max(float(r["rss_mib"]) for r in rows) > 5  # the child really was sampled


def test_download_cli_lists_supported_repos_and_fails_cleanly_offline(tmp_path, capsys):
    assert download_weights.main([]) == 0  # no MODEL: prints the supported repo ids (reference behaviour)
    assert "meta-llama/Meta-Llama-3-8B" in capsys.readouterr().out
    with pytest.raises(Exception):  # no network / no huggingface_hub access: an error, never a silent success
        download_weights.main(["TinyLlama/TinyLlama-1.1B-Chat-v1.0", "--ckpt-dir", str(tmp_path)])
    assert not (tmp_path / "TinyLlama" / "TinyLlama-1.1B-Chat-v1.0" / "lit_model.pth").exists()


def test_owt_shard_writer(tmp_path):
    docs = [list(range(i, i + 7)) for i in range(50)]
    total = sum(len(d) for d in docs)
    prepare_owt.write_sharded(docs, total, tmp_path / "train.bin", n_shards=8)
    arr = np.memmap(tmp_path / "train.bin", dtype=np.uint16, mode="r")
    assert arr.shape == (total,) and arr[:7].tolist() == list(range(7)) and arr[-1] == 49 + 6


def test_optional_import_helpers():
    assert module_available("json") and not module_available("definitely_not_a_module_xyz")
    assert bool(RequirementCache("torch")) and not bool(RequirementCache("definitely-not-a-package-xyz"))
    assert bool(ModuleAvailableCache("math"))
    lazy = LazyModule("json")
    assert lazy.dumps({"a": 1}) == '{"a": 1}'

    @requires("definitely_not_a_module_xyz")
    def f():
        return 1

    with pytest.raises(ModuleNotFoundError):
        f()

    @requires("json")
    def g():
        return 2

    assert g() == 2
