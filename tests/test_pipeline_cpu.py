"""Recurrent-pipeline correctness on CPU: N nodes on 127.0.0.1 (separate GPTServer objects with
their own HTTP control ports and TCP data sockets) must reproduce single-device greedy decode
token for token (SURVEY §4 items b, f, g, h; BASELINE config #1 for GPT-2)."""
import threading
import time

import pytest
import torch

from mdi_llm_b200.models.gpt import GPT
from mdi_llm_b200.models.partition import split_and_store
from mdi_llm_b200.parallel.distributed import GPTDistributed
from mdi_llm_b200.parallel.scheduler import (EagerStageRunner, SamplingParams, secondary_loop, starter_loop)
from mdi_llm_b200.parallel.transport import ChaosPolicy, ring
from mdi_llm_b200.models.stage import build_stage
from mdi_llm_b200.models.partition import split_parameters
from mdi_llm_b200.utils.checkpoint import load_from_pt, random_state_dict, write_random_checkpoint


def _reference_tokens(ck, prompts, n_new):
    cfg, sd = load_from_pt(ck)
    m = GPT(cfg)
    m.load_state_dict(sd, strict=not cfg.tie_embeddings)
    m.eval()
    out = []
    for p in prompts:
        m.clear_kv_cache()
        out.append(m.generate(p, len(p) + n_new, temperature=0.0, top_p=0.0).tolist())
    return out


def _run_cluster(ck, topo, n_nodes, prompts, n_new, pre_split=True, **starter_kw):
    if pre_split and n_nodes > 1:
        cfg, sd = load_from_pt(ck)
        split_and_store(sd, n_nodes, ck, config=cfg)
    secs = [GPTDistributed(f"secondary:{i}", topo, ckpt_dir=ck, dtype="float32") for i in range(n_nodes - 1)]
    st = GPTDistributed("starter", topo, ckpt_dir=ck, dtype="float32", sampling=SamplingParams.greedy(), **starter_kw)
    tok_time = st.start(n_samples=len(prompts), tokens_per_sample=n_new, prompt=prompts, quiet=True)
    for s in secs:  # PUT /stop was sent by the starter: servers wind down on their own
        s.gpt_serv.shutdown()
    return st, tok_time


@pytest.mark.parametrize("family,n_nodes,n_samples", [("llama", 1, 2), ("llama", 2, 2), ("llama", 3, 4), ("gpt2", 2, 2), ("llama", 3, 1)])
def test_loopback_pipeline_token_exact(tmp_path, topology, tiny_llama_cfg, tiny_gpt2_cfg, family, n_nodes, n_samples):
    cfg = tiny_llama_cfg if family == "llama" else tiny_gpt2_cfg
    ck = write_random_checkpoint(tmp_path / "custom" / f"tiny-{family}", cfg, dtype=torch.float32)
    prompts = [torch.tensor([256, 10 + i, 20, 30 + i][: 3 + i % 2]) for i in range(n_samples)]
    n_new = 7
    if n_samples < n_nodes:
        with pytest.warns(UserWarning):
            st, tok_time = _run_cluster(ck, topology(n_nodes), n_nodes, prompts, n_new)
    else:
        st, tok_time = _run_cluster(ck, topology(n_nodes), n_nodes, prompts, n_new)
    res = st.gpt_serv.last_result
    ref = _reference_tokens(ck, prompts, n_new)
    for i in range(n_samples):
        assert res.samples[i].tolist() == ref[i], f"sample {i}"
        assert res.samples[i].size(1) == len(prompts[i]) + n_new  # exactly max_new_tokens generated
    assert tok_time[-1][0] == n_samples * n_new and tok_time[0] == (0, 0.0)
    assert len(st.out_text) == n_samples


def test_chunks_pushed_over_http_when_not_presplit(tmp_path, topology, tiny_llama_cfg):
    """Secondaries without a chunk on disk get their weights inside POST /init (gptserver.py:1160-1169)."""
    ck = write_random_checkpoint(tmp_path / "custom" / "tiny", tiny_llama_cfg, dtype=torch.float32)
    topo = topology(2)
    sec = GPTDistributed("secondary:0", topo, chunk_path=tmp_path / "nowhere.pth", dtype="float32")
    st = GPTDistributed("starter", topo, ckpt_dir=ck, dtype="float32", sampling=SamplingParams.greedy(), push_chunks=True)
    assert not st.model_was_split
    prompts = [torch.tensor([256, 5, 6]), torch.tensor([256, 7, 8])]
    st.start(n_samples=2, tokens_per_sample=5, prompt=prompts, quiet=True)
    sec.gpt_serv.shutdown()
    ref = _reference_tokens(ck, prompts, 5)
    assert [st.gpt_serv.last_result.samples[i].tolist() for i in range(2)] == ref


def test_sequence_length_truncation_and_overflow(tmp_path, topology, tiny_llama_cfg):
    ck = write_random_checkpoint(tmp_path / "custom" / "tiny", tiny_llama_cfg, dtype=torch.float32)
    with pytest.raises(ValueError):
        GPTDistributed("starter", topology(1), ckpt_dir=ck, dtype="float32", model_seq_length=10_000, start_http=False)
    st = GPTDistributed("starter", topology(1), ckpt_dir=ck, dtype="float32", model_seq_length=16,
                        sampling=SamplingParams.greedy())
    assert st.gpt_serv.model.max_seq_length == 16 and st.gpt_serv.model.cos.size(0) == 16
    with pytest.raises(ValueError, match="exceed block size"):
        st.start(n_samples=1, tokens_per_sample=20, prompt=[torch.tensor([1, 2, 3])], quiet=True)


def test_text_prompt_path_with_byte_tokenizer(tmp_path, topology, tiny_llama_cfg):
    ck = write_random_checkpoint(tmp_path / "custom" / "tiny", tiny_llama_cfg, dtype=torch.float32)
    with pytest.warns(UserWarning, match="byte-level"):
        st = GPTDistributed("starter", topology(1), ckpt_dir=ck, dtype="float32", sampling=SamplingParams.greedy())
    st.start(n_samples=2, tokens_per_sample=4, prompt="Hi", quiet=True)
    res = st.gpt_serv.last_result
    assert res.prompt_lengths == {0: 3, 1: 3}  # <bos> H i
    assert len(st.out_text) == 2 and all(isinstance(t, str) for t in st.out_text)


def _inproc_ring(cfg, n_nodes, chaos=None):
    sd = random_state_dict(cfg, dtype=torch.float32)
    chunks, info = split_parameters(dict(sd), n_nodes)
    runners = []
    st = build_stage(cfg, "starter", info["plan"][0], meta=True)
    st.load_weights(chunks["starter"])
    runners.append(EagerStageRunner(st))
    for i, c in enumerate(chunks["secondary"]):
        s = build_stage(cfg, f"secondary:{i}", info["plan"][i + 1], meta=True)
        s.load_weights(c)
        runners.append(EagerStageRunner(s))
    ts = ring(n_nodes)
    if chaos is not None:
        ts[1].chaos = chaos
    return sd, runners, ts


def test_inproc_ring_with_delay_chaos_is_still_exact(tiny_llama_cfg):
    chaos = ChaosPolicy(p_delay=0.5, delay_s=0.002, seed=1)
    sd, runners, ts = _inproc_ring(tiny_llama_cfg, 3, chaos)
    running = threading.Event()
    running.set()
    workers = [threading.Thread(target=secondary_loop, args=(runners[i], ts[i], running), kwargs={"recv_timeout": 0.05}, daemon=True)
               for i in (1, 2)]
    for w in workers:
        w.start()
    prompts = [torch.tensor([1, 2, 3]), torch.tensor([4, 5, 6, 7]), torch.tensor([8, 9])]
    res = starter_loop(runners[0], ts[0], prompts, 6, SamplingParams.greedy(), running, n_nodes=3)
    running.clear()
    for w in workers:
        w.join(timeout=2)
    m = GPT(tiny_llama_cfg)
    m.load_state_dict(sd)
    m.eval()
    for i, p in enumerate(prompts):
        m.clear_kv_cache()
        assert res.samples[i].tolist() == m.generate(p, len(p) + 6, temperature=0.0, top_p=0.0).tolist()
    assert chaos.delayed > 0


def test_watchdog_fires_when_a_hop_is_dropped(tiny_llama_cfg):
    """A dead/dropping node hangs the reference's ring forever; here the starter's
    watchdog turns it into an error."""
    chaos = ChaosPolicy(p_drop=1.0)
    _, runners, ts = _inproc_ring(tiny_llama_cfg, 2, chaos)
    running = threading.Event()
    running.set()
    w = threading.Thread(target=secondary_loop, args=(runners[1], ts[1], running), kwargs={"recv_timeout": 0.05}, daemon=True)
    w.start()
    with pytest.raises(TimeoutError):
        starter_loop(runners[0], ts[0], [torch.tensor([1, 2, 3])], 4, SamplingParams.greedy(), running,
                     n_nodes=2, recv_timeout=0.05, watchdog_s=0.3)
    assert not running.is_set()
    w.join(timeout=2)


# ---- earlier protocol generations ----------------------------------------------------
def _greedy_ref(cfg, sd, prompts, n_new):
    m = GPT(cfg)
    m.load_state_dict(sd, strict=not cfg.tie_embeddings)
    m.eval()
    out = []
    for p in prompts:
        m.clear_kv_cache()
        out.append(m.generate(p, len(p) + n_new, temperature=0.0, top_p=0.0).tolist())
    return out


def test_cacheless_context_resend_mode_matches_cached_decode(tiny_gpt2_cfg):
    """GPT-2 generation protocol: no KV caches, the whole (growing) context travels the ring each step
    (old/GPT2/sub/model_dist.py:959-972).  Same tokens as cached decode while the context fits."""
    sd, runners, ts = _inproc_ring(tiny_gpt2_cfg, 2)
    running = threading.Event()
    running.set()
    w = threading.Thread(target=secondary_loop, args=(runners[1], ts[1], running),
                         kwargs={"recv_timeout": 0.05, "use_kv_cache": False}, daemon=True)
    w.start()
    prompts = [torch.tensor([1, 2, 3]), torch.tensor([4, 5])]
    res = starter_loop(runners[0], ts[0], prompts, 6, SamplingParams.greedy(), running, n_nodes=2,
                       use_kv_cache=False, block_size=tiny_gpt2_cfg.block_size)
    running.clear()
    w.join(timeout=2)
    assert runners[0].model.kv_pool is None and runners[1].model.kv_pool is None  # really cache-less
    assert [res.samples[i].tolist() for i in range(2)] == _greedy_ref(tiny_gpt2_cfg, sd, prompts, 6)


def test_cacheless_mode_crops_to_block_size(tiny_llama_cfg):
    sd, runners, ts = _inproc_ring(tiny_llama_cfg, 2)
    running = threading.Event()
    running.set()
    w = threading.Thread(target=secondary_loop, args=(runners[1], ts[1], running),
                         kwargs={"recv_timeout": 0.05, "use_kv_cache": False}, daemon=True)
    w.start()
    res = starter_loop(runners[0], ts[0], [torch.tensor([1, 2, 3, 4, 5, 6])], 5, SamplingParams.greedy(), running,
                       n_nodes=2, use_kv_cache=False, block_size=8)
    running.clear()
    w.join(timeout=2)
    assert res.samples[0].size(1) == 11  # longer than the 8-token window: the context slid


@pytest.mark.parametrize("n_nodes", [2, 3])
def test_finisher_topology_token_exact(tmp_path, topology, tiny_llama_cfg, n_nodes):
    """First-generation chain starter -> intermediate -> finisher: the last node owns ln_f + lm_head
    and returns logits (old/nanoGPT/sub/model_dist.py:90-221)."""
    ck = write_random_checkpoint(tmp_path / "custom" / "tiny-fin", tiny_llama_cfg, dtype=torch.float32)
    topo = topology(n_nodes)
    prompts = [torch.tensor([256, 10 + i, 20]) for i in range(n_nodes)]
    secs = [GPTDistributed(f"secondary:{i}", topo, ckpt_dir=ck, dtype="float32") for i in range(n_nodes - 1)]
    st = GPTDistributed("starter", topo, ckpt_dir=ck, dtype="float32", sampling=SamplingParams.greedy(),
                        head_on="finisher")
    st.start(n_samples=len(prompts), tokens_per_sample=6, prompt=prompts, quiet=True)
    for s in secs:
        s.gpt_serv.shutdown()
    assert not hasattr(st.gpt_serv.model, "lm_head")
    assert secs[-1].gpt_serv.model.role == "finisher"
    assert [st.gpt_serv.last_result.samples[i].tolist() for i in range(len(prompts))] == _reference_tokens(ck, prompts, 6)


def test_cacheless_mode_through_gptdistributed(tmp_path, topology, tiny_gpt2_cfg):
    ck = write_random_checkpoint(tmp_path / "custom" / "tiny-gpt2-nc", tiny_gpt2_cfg, dtype=torch.float32)
    topo = topology(2)
    prompts = [torch.tensor([256, 3, 4]), torch.tensor([256, 9])]
    sec = GPTDistributed("secondary:0", topo, ckpt_dir=ck, dtype="float32")
    st = GPTDistributed("starter", topo, ckpt_dir=ck, dtype="float32", sampling=SamplingParams.greedy(),
                        use_kv_cache=False)
    st.start(n_samples=2, tokens_per_sample=5, prompt=prompts, quiet=True)
    sec.gpt_serv.shutdown()
    assert sec.gpt_serv.use_kv_cache is False
    assert [st.gpt_serv.last_result.samples[i].tolist() for i in range(2)] == _reference_tokens(ck, prompts, 5)


@pytest.mark.parametrize("policy,n_nodes", [("half", 3), ("third", 3), ("third", 4)])
def test_sub_layer_partitions_token_exact(tmp_path, topology, tiny_llama_cfg, policy, n_nodes):
    """Stage boundaries INSIDE layers (attention | gate/up | down units; the boundary after a gate/up unit carries
    ``[x | h]``): chunks split on the fly by ``GPTDistributed(partition=...)``, every node infers its shape from
    its chunk file, and the ring reproduces single-device greedy decode token for token."""
    ck = write_random_checkpoint(tmp_path / "custom" / f"tiny-{policy}", tiny_llama_cfg, dtype=torch.float32)
    prompts = [torch.tensor([256, 10 + i, 20, 30 + i][: 3 + i % 2]) for i in range(n_nodes + 1)]
    st, _ = _run_cluster(ck, topology(n_nodes), n_nodes, prompts, 6, pre_split=False, partition=policy)
    from mdi_llm_b200.models.partition import stage_shape_from_state_dict
    from mdi_llm_b200.utils.checkpoint import lazy_load

    shapes = [stage_shape_from_state_dict(lazy_load(ck / "chunks" / f"{n_nodes}nodes" / f))
              for f in ["model_starter.pth"] + [f"model_secondary{i}.pth" for i in range(n_nodes - 1)]]
    assert any(s["first_parts"] != "both" or s["last_parts"] != "both" for s in shapes)  # a layer really was cut
    if policy == "third":
        assert any(s["last_parts"] == "attn_gu" or s["first_parts"] == "down" for s in shapes) or n_nodes < 4
    ref = _reference_tokens(ck, prompts, 6)
    for i in range(len(prompts)):
        assert st.gpt_serv.last_result.samples[i].tolist() == ref[i], f"sample {i}"
