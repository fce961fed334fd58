"""Property-based tests (hypothesis): partition planners against brute force, splitters as exact partitions of the state
dict, the restricted unpickler and the wire format under generated inputs."""
import itertools
import pickle

import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from mdi_llm_b200.models.config import Config
from mdi_llm_b200.models.partition import (_plan_units, half_stages, plan_half_units, plan_third_units, split_parameters_half,
                                           split_parameters_units, stage_shape_from_state_dict, stage_specs, third_stages)
from mdi_llm_b200.parallel.transport.socket_transport import encode_frame
from mdi_llm_b200.utils.checkpoint import random_state_dict
from mdi_llm_b200.utils.safe_pickle import UnsafePayload, safe_loads

FAST = settings(max_examples=60, deadline=None, derandomize=True, database=None,
                suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])  # same examples on every run


def _brute_force(costs, n, head, fixed):
    U, best = len(costs), float("inf")
    for cuts in itertools.combinations(range(1, U), n - 1):
        b = (0,) + cuts + (U,)
        worst = max(sum(costs[b[i]:b[i + 1]]) + fixed + (head if i == 0 else 0.0) for i in range(n))
        best = min(best, worst)
    return best


@FAST
@given(costs=st.lists(st.floats(0.5, 50.0), min_size=2, max_size=9), n=st.integers(1, 4), head=st.floats(0.0, 80.0),
       fixed=st.floats(0.0, 10.0))
def test_unit_planner_is_optimal(costs, n, head, fixed):
    """The DP returns a contiguous partition whose slowest stage equals the brute-force optimum."""
    if n > len(costs):
        with pytest.raises(ValueError):
            _plan_units(costs, n, head, fixed)
        return
    plan = _plan_units(costs, n, head, fixed)
    assert len(plan) == n and sum(plan) == len(costs) and min(plan) >= 1
    b = [0] + list(itertools.accumulate(plan))
    worst = max(sum(costs[b[i]:b[i + 1]]) + fixed + (head if i == 0 else 0.0) for i in range(n))
    assert worst <= _brute_force(costs, n, head, fixed) + 1e-6


def _cfg(n_layer, moe=False, plain=False):
    kw = dict(n_layer=n_layer, n_embd=32, n_head=4, n_query_groups=2, intermediate_size=48, vocab_size=90, padded_vocab_size=96,
              block_size=32)
    if moe:
        kw.update(mlp_class_name="LLaMAMoE", n_expert=3, n_expert_per_token=2)
    if plain:
        kw.update(mlp_class_name="GptNeoxMLP", norm_class_name="LayerNorm", bias=True)
    return Config.from_name("tiny-llama-1.1b", **kw)


@FAST
@given(n_layer=st.integers(1, 6), n_nodes=st.integers(1, 5), policy=st.sampled_from(["third", "half", "balanced", "auto"]),
       arch=st.sampled_from(["gated", "moe", "plain"]))
def test_stage_specs_tile_the_model(n_layer, n_nodes, policy, arch):
    """Whatever the policy and MLP family: the stages are contiguous, cover every sub-layer exactly once, start where the
    previous one stopped, and every stage owns at least one unit."""
    cfg = _cfg(n_layer, moe=arch == "moe", plain=arch == "plain")
    if n_nodes > n_layer:  # whole-layer policies cannot give every node a layer; sub-layer ones may
        try:
            specs = stage_specs(n_nodes, cfg, policy)
        except ValueError:
            return
    else:
        specs = stage_specs(n_nodes, cfg, policy)
    assert len(specs) == n_nodes
    order = {"attn": 0, "gu": 1, "down": 2}
    from mdi_llm_b200.models.gpt import part_units

    seq = []
    for sp in specs:
        nb = sp["n_blocks"]
        assert nb >= 1
        for j in range(nb):
            if nb == 1:
                units = [u for u in part_units(sp["first_parts"]) if u in part_units(sp["last_parts"])]
            elif j == 0:
                units = part_units(sp["first_parts"])
            elif j == nb - 1:
                units = part_units(sp["last_parts"])
            else:
                units = ("attn", "gu", "down")
            seq += [3 * (sp["layer_offset"] + j) + order[u] for u in units]
    assert seq == list(range(3 * n_layer)), (specs, seq)


@FAST
@given(n_layer=st.integers(1, 4), data=st.data())
def test_third_and_half_splits_partition_the_state_dict(n_layer, data):
    """Every tensor lands in exactly one chunk (bit-identical), and the chunk's keys alone tell its stage shape."""
    cfg = _cfg(n_layer)
    n_nodes = data.draw(st.integers(1, min(4, 3 * n_layer)))
    units = data.draw(st.lists(st.integers(1, 3 * n_layer), min_size=n_nodes, max_size=n_nodes).filter(lambda u: sum(u) == 3 * n_layer))
    sd = random_state_dict(cfg, dtype=torch.float32, seed=1, std=0.02)
    ref = {k: v.clone() for k, v in sd.items()}
    specs = third_stages(units)
    chunks = split_parameters_units(dict(sd), specs)
    seen = {}
    for sp, ch in zip(specs, [chunks["starter"]] + list(chunks["secondary"])):
        shape = stage_shape_from_state_dict(ch)
        assert (shape["n_blocks"], shape["first_parts"], shape["last_parts"]) == (sp["n_blocks"], sp["first_parts"], sp["last_parts"])
        for k, v in ch.items():
            if k.startswith("transformer.h."):
                _, _, li, tail = k.split(".", 3)
                k = f"transformer.h.{int(li) + sp['layer_offset']}.{tail}"
            assert k not in seen
            seen[k] = v
    assert seen.keys() == ref.keys() and all(torch.equal(seen[k], ref[k]) for k in ref)


@FAST
@given(n_layer=st.integers(1, 5), n_nodes=st.integers(2, 5))
def test_half_plan_and_split_agree(n_layer, n_nodes):
    cfg = _cfg(n_layer)
    if n_nodes > 2 * n_layer:
        with pytest.raises(ValueError):
            plan_half_units(n_nodes, cfg)
        return
    units = plan_half_units(n_nodes, cfg)
    assert sum(units) == 2 * n_layer and min(units) >= 1
    sd = random_state_dict(cfg, dtype=torch.float32, seed=2, std=0.02)
    n_keys = len(sd)
    chunks = split_parameters_half(dict(sd), units)
    parts = [chunks["starter"]] + list(chunks["secondary"])
    assert sum(len(p) for p in parts) == n_keys
    for hs, ch in zip(half_stages(units), parts):
        shape = stage_shape_from_state_dict(ch)
        assert (shape["n_blocks"], shape["first_mlp_only"], shape["last_attn_only"]) == (hs.n_blocks, hs.first_mlp_only, hs.last_attn_only)
    if cfg.mlp_class_name in ("LLaMAMLP", "GemmaMLP") and n_nodes <= 3 * n_layer:
        t = plan_third_units(n_nodes, cfg)
        assert sum(t) == 3 * n_layer and min(t) >= 1


_leaf = st.one_of(st.none(), st.booleans(), st.integers(-2 ** 40, 2 ** 40), st.floats(allow_nan=False), st.text(max_size=20),
                  st.binary(max_size=20))
_tree = st.recursive(_leaf, lambda ch: st.one_of(st.lists(ch, max_size=4), st.tuples(ch, ch),
                                                 st.dictionaries(st.text(max_size=8), ch, max_size=4)), max_leaves=12)


@FAST
@given(obj=_tree, shape=st.lists(st.integers(0, 5), min_size=0, max_size=3), dtype=st.sampled_from([torch.float32, torch.bfloat16, torch.int64]))
def test_safe_loads_roundtrips_plain_messages_with_tensors(obj, shape, dtype):
    t = (torch.arange(int(torch.tensor(shape).prod()) if shape else 1).reshape(shape or [1]) % 7).to(dtype)
    msg = {"payload": obj, "data": t, "nested": [t[:0], {"k": (obj, 1)}]}
    back = safe_loads(pickle.dumps(msg))
    assert back["payload"] == obj and torch.equal(back["data"], t) and back["data"].dtype == dtype
    assert back["nested"][1]["k"][0] == obj


class _Evil:
    def __init__(self, fn, *args):
        self.fn, self.args = fn, args

    def __reduce__(self):
        return self.fn, self.args


@FAST
@given(which=st.sampled_from(["system", "eval", "exec", "open", "getattr", "import", "popen"]), wrap=st.integers(0, 3))
def test_safe_loads_rejects_callables_wherever_they_hide(which, wrap):
    import importlib
    import os
    import subprocess

    fn, args = {"system": (os.system, ("true",)), "eval": (eval, ("1+1",)), "exec": (exec, ("x=1",)), "open": (open, ("/etc/passwd",)),
                "getattr": (getattr, (_Evil(importlib.import_module, "builtins"), "eval")), "import": (importlib.import_module, ("os",)),
                "popen": (subprocess.Popen, (["true"],))}[which]
    obj = _Evil(fn, *args)
    for i in range(wrap):
        obj = [{"k": obj}, (obj,), {"data": torch.zeros(1), "x": obj}][i % 3]
    with pytest.raises(UnsafePayload):
        safe_loads(pickle.dumps({"sample_index": 0, "data": obj, "stop": False}))


@FAST
@given(idx=st.integers(0, 10 ** 6), stop=st.booleans(), n=st.integers(0, 64))
def test_frame_header_is_the_reference_format(idx, stop, n):
    """16 ASCII characters, left-aligned decimal length, then exactly that many bytes of pickle (connections.py:325-342)."""
    msg = {"sample_index": idx, "data": torch.arange(n, dtype=torch.float32), "stop": stop}
    frame = encode_frame(msg)
    header, body = frame[:16], frame[16:]
    assert len(header) == 16 and header.decode("ascii").rstrip(" ").isdigit() and header.decode("ascii")[0] != " "
    assert int(header.decode("ascii")) == len(body)
    back = safe_loads(body)
    assert back["sample_index"] == idx and back["stop"] == stop and torch.equal(back["data"], msg["data"])


@FAST
@given(vals=st.lists(st.integers(-128, 128), min_size=2, max_size=40, unique=True).map(lambda v: [x / 16 for x in v]),
       k=st.integers(1, 40), top_p=st.floats(0.05, 1.0),
       temp=st.floats(0.1, 2.0), seed=st.integers(0, 1000))
def test_sampler_only_draws_from_the_allowed_set(vals, k, top_p, temp, seed):
    """top-k then nucleus (model.py:42-90): the drawn token is always among the k largest logits AND inside the smallest
    prefix of them whose probability mass reaches top_p; temperature 0 with top_p 0 is the arg-max."""
    from mdi_llm_b200.models.gpt import sample

    logits = torch.tensor(vals, dtype=torch.float32).view(1, 1, -1)
    g = torch.Generator().manual_seed(seed)
    tok = int(sample(logits, temperature=temp, top_k=k, top_p=top_p, generator=g))
    order = sorted(range(len(vals)), key=lambda i: -vals[i])
    kept = order[: min(k, len(vals))]
    assert tok in kept
    probs = torch.tensor([vals[i] for i in kept]).div(temp).softmax(-1)
    mass_before = float(probs[: kept.index(tok)].sum())  # mass of the strictly more likely kept tokens
    assert mass_before < top_p + 1e-4
    assert int(sample(logits, temperature=0.0, top_k=k, top_p=0.0)) == order[0]


@FAST
@given(text=st.text(alphabet=st.characters(blacklist_categories=("Cs",)), max_size=60))
def test_byte_tokenizer_roundtrips_any_text(tmp_path_factory, text):
    from mdi_llm_b200.text.tokenizer import Tokenizer, write_bytes_tokenizer

    d = tmp_path_factory.mktemp("tok")
    write_bytes_tokenizer(d)
    tok = Tokenizer(d, force_backend="bytes")
    ids = tok.encode(text, bos=False)
    assert tok.decode(ids) == text and int(ids.max() if ids.numel() else 0) < tok.vocab_size


@settings(max_examples=25, deadline=None, derandomize=True, database=None)
@given(corpus=st.text(alphabet="abcdefgh ,.\n", min_size=20, max_size=200), probe=st.text(alphabet="abcdefgh ,.\n", max_size=40),
       merges=st.integers(0, 30))
def test_trained_bpe_and_char_tokenizers_roundtrip(corpus, probe, merges):
    from mdi_llm_b200.text.simple_tokenizers import BPETokenizer, CharacterTokenizer

    bpe = BPETokenizer()
    bpe.tokenize(corpus, out_vocab_size=256 + merges)
    assert bpe.decode(bpe.encode(probe)) == probe
    assert len(bpe.encode(corpus)) <= len(corpus.encode("utf-8"))
    ch = CharacterTokenizer()
    ch.tokenize(corpus + probe)
    assert ch.decode(ch.encode(probe)) == probe


class _ScriptedPipe:
    """Stands in for a one-stage DevicePipeline in host-fed mode: hands out a scripted token per decode round."""

    def __init__(self, script):
        self.script, self.i = list(script), 0

    def prepare(self, prompts, max_new):
        self.i = 0

    def prefill(self):
        pass

    def decode_rounds_host(self, n, on_token=None):
        for _ in range(n):
            if self.i < len(self.script):
                on_token(0, self.i, self.script[self.i])
                self.i += 1
        return n, 0, 0


def _litgpt_chat_oracle(produced, stops):
    """The buffering rule of the reference's ``generate_chat`` (model.py:556-573), transcribed over a finished token list."""
    out, yield_i, buf = [], 0, max((len(s) for s in stops), default=1)
    for t in range(1, len(produced) + 1):
        if any(len(s) <= t and tuple(produced[t - len(s):t]) == tuple(s) for s in stops):
            return out, t
        if t - yield_i >= buf:
            out += produced[yield_i:t]
            yield_i = t
    return out + produced[yield_i:], None


@FAST
@given(script=st.lists(st.integers(0, 4), min_size=0, max_size=24), stops=st.lists(st.lists(st.integers(0, 4), min_size=1, max_size=3), max_size=3),
       budget=st.integers(1, 30))
def test_chat_stream_follows_the_reference_buffering_rule(script, stops, budget):
    """`cli/chat.py::stream_device` (the fused-engine twin of `generate_chat`): yields what the reference's rule yields for
    the same token stream — a prefix of the generated tokens that stops before the token completing a stop sequence."""
    from mdi_llm_b200.cli.chat import stream_device

    stops_t = tuple(tuple(s) for s in stops)
    got = list(stream_device(_ScriptedPipe(script), torch.tensor([1, 2, 3]), budget, stops_t))
    produced = script[:budget]
    expect, fired = _litgpt_chat_oracle(produced, stops_t)
    assert got == expect and got == produced[:len(got)]
    if fired is None:
        assert got == produced
    else:
        assert len(got) < fired  # the completing token is never printed
