import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "multigpu: test needs >= 2 CUDA devices")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        n = torch.cuda.device_count()
        skip_multi = pytest.mark.skip(reason="needs >= 2 GPUs")
        for item in items:
            if "multigpu" in item.keywords and n < 2:
                item.add_marker(skip_multi)
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def free_ports(n):
    socks, ports = [], []
    for _ in range(n):
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        socks.append(s)
        ports.append(s.getsockname()[1])
    for s in socks:
        s.close()
    return ports


@pytest.fixture
def topology():
    """Factory for a loopback node-topology dict in the reference's JSON schema."""

    def make(n_nodes, device="cpu"):
        ports = free_ports(3 * n_nodes)

        def node(i):
            return {"addr": "127.0.0.1",
                    "communication": {"port": ports[3 * i], "starter_addr": "127.0.0.1"},
                    "inference": {"port_in": ports[3 * i + 1], "port_out": ports[3 * i + 2]},
                    "device": device}

        return {"nodes": {"starter": node(0), "secondary": [node(i) for i in range(1, n_nodes)]}}

    return make


@pytest.fixture
def tiny_llama_cfg():
    from mdi_llm_b200.models.config import Config

    return Config.from_name("tiny-llama-1.1b", n_layer=5, n_embd=64, n_head=4, n_query_groups=2,
                            intermediate_size=128, vocab_size=300, padded_vocab_size=320, block_size=64)


@pytest.fixture
def tiny_gpt2_cfg():
    from mdi_llm_b200.models.config import Config

    return Config.from_name("gpt2", n_layer=5, n_embd=48, n_head=4, block_size=64, vocab_size=300,
                            padded_vocab_size=300)
