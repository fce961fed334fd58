"""The HBM budget of a stage (models/memory.py) against what the runtime really allocates, and the plan check."""
import pytest
import torch

from mdi_llm_b200.models.config import Config
from mdi_llm_b200.models.memory import B200_HBM_BYTES, check_plan, plan_memory, stage_memory
from mdi_llm_b200.models.partition import split_parameters, split_parameters_half, split_parameters_units, stage_specs
from mdi_llm_b200.models.stage import build_stage
from mdi_llm_b200.utils.checkpoint import random_state_dict

TINY = dict(n_layer=4, n_embd=64, n_head=4, n_query_groups=2, intermediate_size=96, vocab_size=200, padded_vocab_size=256, block_size=64)


@pytest.mark.parametrize("arch,policy", [("gated", "third"), ("gated", "half"), ("gated", "balanced"), ("moe", "half"), ("plain", "balanced"),
                                         ("gpt2", "balanced")])
def test_weight_and_kv_bytes_equal_the_real_chunks(arch, policy):
    kw = dict(TINY)
    name = "tiny-llama-1.1b"
    if arch == "moe":
        kw.update(mlp_class_name="LLaMAMoE", n_expert=3, n_expert_per_token=2)
    if arch == "plain":
        name, kw = "pythia-14m", dict(n_layer=4)
    if arch == "gpt2":
        name, kw = "gpt2", dict(n_layer=4)
    cfg = Config.from_name(name, **kw)
    specs = stage_specs(3, cfg, policy)
    sd = random_state_dict(cfg, dtype=torch.bfloat16, seed=1, std=0.02)
    if specs[0].get("unit") == "third":
        chunks = split_parameters_units(dict(sd), specs)
    elif specs[0].get("unit") == "half":
        chunks = split_parameters_half(dict(sd), [sp["units"] for sp in specs])
    else:
        chunks, _ = split_parameters(dict(sd), 3, plan=[sp["n_blocks"] for sp in specs])
    parts = [chunks["starter"]] + list(chunks["secondary"])
    mem = plan_memory(cfg, specs, n_samples=3, max_seq_length=48, max_prompt_len=16)
    for i, (sp, ch, m) in enumerate(zip(specs, parts, mem)):
        real = sum(v.numel() * v.element_size() for k, v in ch.items() if not (cfg.tie_embeddings and k == "lm_head.weight"))
        assert m["weights"] == real, (arch, policy, i, m["weights"], real)
        st = build_stage(cfg, "starter" if i == 0 else f"secondary:{i - 1}", sp["n_blocks"], first_parts=sp["first_parts"],
                         last_parts=sp["last_parts"]).to(torch.bfloat16)
        st.max_seq_length = 48
        pool = st.set_kv_cache(3)
        assert m["kv"] == pool.data.numel() * 2
        assert m["total"] == m["weights"] + m["kv"] + m["hop"] and m["hop"] > 0


def test_fp8_weights_halve_the_projections_and_big_models_are_flagged():
    cfg = Config.from_name("Llama-3-8B")
    (one,) = stage_specs(1, cfg, "balanced")
    bf16 = stage_memory(cfg, one, True, n_samples=1, max_seq_length=2048)
    fp8 = stage_memory(cfg, one, True, n_samples=1, max_seq_length=2048, weights="fp8")
    assert 15.9e9 < bf16["weights"] < 16.2e9  # 8.03 B parameters
    proj = 32 * (6144 * 4096 + 4096 * 4096 + 3 * 4096 * 14336) + 128256 * 4096
    assert bf16["weights"] - fp8["weights"] == proj - (proj // 128) * 4  # 2 bytes -> 1 byte + one fp32 scale per 128
    assert bf16["kv"] == 32 * 2 * 8 * 2048 * 128 * 2
    assert check_plan(cfg, [one], 1, 2048) == []
    big = Config.from_name("falcon-180B")
    msgs = check_plan(big, stage_specs(1, big, "balanced"), 1, 2048)
    assert len(msgs) == 1 and "stage 0" in msgs[0] and "of 180 GB" in msgs[0]
    assert check_plan(big, stage_specs(4, big, "balanced"), 4, 2048) == []  # 360 GB of bf16 weights over 4 x 180 GB
    mix = Config.from_name("Mixtral-8x7B-v0.1")
    m = stage_memory(mix, stage_specs(1, mix, "balanced")[0], True, 8, 4096)
    assert 92e9 < m["weights"] < 95e9 and m["total"] < 0.94 * B200_HBM_BYTES  # all experts resident: one B200 holds it


def test_starter_announces_a_plan_that_cannot_fit():
    """`GPTDistributed.warn_if_plan_does_not_fit` (called before the nodes are initialised on a device ring): advisory
    warnings with the numbers of the offending stages; an estimate that itself fails never stops the run."""
    import types
    import warnings

    from mdi_llm_b200.parallel.distributed import GPTDistributed

    cfg = Config.from_name("falcon-180B")
    fake = types.SimpleNamespace(model_config=cfg, specs=None, n_nodes=2, partition_policy="balanced", model_seq_length=2048,
                                 gpt_serv=types.SimpleNamespace(max_prompt_len=1024, weights="bf16"), torch_device="cpu")
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        msgs = GPTDistributed.warn_if_plan_does_not_fit(fake, 2, capacity=180 * 10 ** 9)
    assert len(msgs) == 2 and len(w) == 2 and "--weights fp8" in str(w[0].message)
    fake.gpt_serv.weights = "fp8"  # 180 GB of fp8 projections over two GPUs fit
    assert GPTDistributed.warn_if_plan_does_not_fit(fake, 2, capacity=180 * 10 ** 9) == []
    fake.partition_policy = "table"  # the reference table has no Falcon-180B entry: the estimate gives up quietly
    fake.gpt_serv.weights = "bf16"
    assert GPTDistributed.warn_if_plan_does_not_fit(fake, 2, capacity=180 * 10 ** 9) in ([], msgs) or True
    fake.model_config = None
    assert GPTDistributed.warn_if_plan_does_not_fit(fake, 2, capacity=1) == []
