import torch
import pytest

from mdi_llm_b200.models.config import Config
from mdi_llm_b200.models.partition import (N_LAYERS_NODES, balanced_plan, count_transformer_blocks,
                                            merge_chunks, plan_layers, split_and_store, split_parameters)
from mdi_llm_b200.utils.checkpoint import (incremental_save, init_from_state_dict, lazy_load, load_from_pt,
                                           random_state_dict, write_random_checkpoint)


def test_reference_table_values():
    assert N_LAYERS_NODES[2][32] == {"N_LAYERS_START": 14, "N_LAYERS_SECONDARY": 18}
    assert N_LAYERS_NODES[3][22] == {"N_LAYERS_START": 6, "N_LAYERS_SECONDARY": 8}
    assert N_LAYERS_NODES[1][12] == {"N_LAYERS_START": 12}
    assert plan_layers(4, 32) == [5, 9, 9, 9] and plan_layers(5, 22) == [2, 5, 5, 5, 5]


def test_planner_for_missing_topologies():
    cfg = Config.from_name("Llama-3-8B")
    plan = plan_layers(8, 32, cfg)
    assert sum(plan) == 32 and len(plan) == 8 and min(plan) >= 1
    head = cfg.head_param_count() / cfg.block_param_count()
    assert max(plan[0] + head, max(plan[1:])) <= 5.0 + 1e-6  # uniform 4/4 would cost 6.4
    with pytest.raises(KeyError):
        plan_layers(8, 32, cfg, policy="table")
    assert balanced_plan(1, 7) == [7]
    with pytest.raises(ValueError):
        balanced_plan(4, 3)


@pytest.mark.parametrize("n_nodes", [2, 3])
def test_split_reindex_and_merge_identity(tiny_llama_cfg, n_nodes):
    sd = random_state_dict(tiny_llama_cfg, dtype=torch.float32)
    ref = {k: v.clone() for k, v in sd.items()}
    chunks, info = split_parameters(sd, n_nodes)
    assert len(sd) == 0  # everything consumed
    plan = info["plan"]
    assert count_transformer_blocks(chunks["starter"]) == plan[0]
    assert "lm_head.weight" in chunks["starter"] and "transformer.ln_f.weight" in chunks["starter"]
    for i, c in enumerate(chunks["secondary"]):
        assert count_transformer_blocks(c) == plan[i + 1]
        assert all(k.startswith("transformer.h.") for k in c)
        assert "transformer.h.0.norm_1.weight" in c  # locally re-indexed from 0
    merged = merge_chunks(chunks["starter"], chunks["secondary"])
    assert merged.keys() == ref.keys()
    assert all(torch.equal(merged[k], ref[k]) for k in ref)


def test_split_and_store_layout(tmp_path, tiny_llama_cfg):
    ck = write_random_checkpoint(tmp_path / "org" / "model", tiny_llama_cfg, dtype=torch.float32)
    cfg, sd = load_from_pt(ck)
    assert cfg.asdict() == tiny_llama_cfg.asdict()
    out = split_and_store(sd, 3, ck)
    assert out == ck / "chunks" / "3nodes"
    assert sorted(p.name for p in out.iterdir()) == ["model_secondary0.pth", "model_secondary1.pth", "model_starter.pth"]
    lazy = lazy_load(out / "model_secondary1.pth")
    assert count_transformer_blocks(lazy) == 2


def test_init_from_state_dict_on_meta_and_ties(tiny_gpt2_cfg):
    from mdi_llm_b200.models.gpt import GPT

    sd = random_state_dict(tiny_gpt2_cfg, dtype=torch.float32)
    with torch.device("meta"):
        m = GPT(tiny_gpt2_cfg)
    init_from_state_dict(m, sd)
    assert all(p.device.type == "cpu" for p in m.parameters())
    assert m.lm_head.weight.data_ptr() == m.transformer.wte.weight.data_ptr()
    with pytest.raises(KeyError):
        with torch.device("meta"):
            m2 = GPT(tiny_gpt2_cfg)
        init_from_state_dict(m2, {k: v for k, v in sd.items() if "ln_f" not in k})


def test_incremental_save_roundtrip(tmp_path):
    a, b = torch.randn(17, 5), torch.arange(12, dtype=torch.int32).reshape(3, 4)
    with incremental_save(tmp_path / "out.pth") as saver:
        sd = {"a": saver.store_early(a), "b": saver.store_early(b)}
        saver.save(sd)
    back = torch.load(tmp_path / "out.pth", weights_only=True)
    assert torch.equal(back["a"], a) and torch.equal(back["b"], b)


def test_fp8_block_quantisation_roundtrip_error():
    from mdi_llm_b200.utils.quantize import dequantize_fp8_block, fp8_error, quantize_fp8_block, quantize_linear_weights

    torch.manual_seed(0)
    w = torch.randn(96, 512) * 0.02
    w[3, 100] = 1.5  # an outlier only hurts its own 128-block
    q, s = quantize_fp8_block(w)
    assert q.dtype == torch.float8_e4m3fn and s.shape == (96, 4) and s.dtype == torch.float32
    d = dequantize_fp8_block(q, s, torch.float32)
    assert fp8_error(w) < 0.04
    clean = torch.ones(96, 512, dtype=torch.bool)
    clean[3, :128] = False
    assert ((d - w).abs()[clean] <= 0.0625 * w.abs()[clean] + 1e-6).all()  # e4m3: 3 mantissa bits
    sd = quantize_linear_weights({"transformer.wte.weight": torch.randn(10, 128), "transformer.h.0.attn.proj.weight": w,
                                  "transformer.h.0.norm_1.weight": torch.ones(512)})
    assert sd["transformer.h.0.attn.proj.weight"].dtype == torch.float8_e4m3fn
    assert "transformer.h.0.attn.proj.weight_scale" in sd and sd["transformer.wte.weight"].dtype == torch.float32
    with pytest.raises(ValueError):
        quantize_fp8_block(torch.randn(4, 100))


# ---- half-layer partitions ------------------------------------------------------------------------------
def test_half_unit_plan_covers_and_balances():
    from mdi_llm_b200.models.config import Config
    from mdi_llm_b200.models.partition import decode_unit_costs, half_stages, plan_half_units, plan_layers

    cfg = Config.from_name("Llama-3-8B")
    ca, cm, ch = decode_unit_costs(cfg)
    for n in (1, 2, 4, 8):
        plan = plan_half_units(n, cfg)
        assert len(plan) == n and sum(plan) == 2 * cfg.n_layer and min(plan) >= 1
        st = half_stages(plan)
        cost = [sum(ca if u % 2 == 0 else cm for u in range(s.lo_unit, s.hi_unit)) + (ch if i == 0 else 0) for i, s in enumerate(st)]
        whole = plan_layers(n, cfg.n_layer, cfg, policy="balanced")
        whole_cost = [l * (ca + cm) + (ch if i == 0 else 0) for i, l in enumerate(whole)]
        assert max(cost) <= max(whole_cost) + 1e-6  # never worse than whole-layer stages
    assert max(cost) < max(whole_cost)  # and strictly better for 8 x Llama-3-8B
    with pytest.raises(ValueError):
        plan_half_units(2, Config.from_name("pythia-14m"))  # parallel residual cannot be cut


def test_half_stage_chain_equals_full_model(tiny_llama_cfg):
    """Stages cut inside layers (attention | MLP) reproduce the full model exactly, with KV caches."""
    import torch
    from mdi_llm_b200.models.gpt import GPT
    from mdi_llm_b200.models.partition import half_stages, split_parameters_half
    from mdi_llm_b200.models.stage import build_stage
    from mdi_llm_b200.utils.checkpoint import random_state_dict

    cfg = tiny_llama_cfg
    sd = random_state_dict(cfg, dtype=torch.float32)
    full = GPT(cfg)
    full.load_state_dict(sd)
    full.eval()
    plan = [3, 2, 2 * cfg.n_layer - 5]  # cuts after unit 3 (inside layer 1) and unit 5 (inside layer 2)
    st = half_stages(plan)
    assert st[0].last_attn_only and st[1].first_mlp_only and st[1].last_attn_only and st[2].first_mlp_only
    chunks = split_parameters_half(dict(sd), plan)
    mods = []
    for i, s in enumerate(st):
        m = build_stage(cfg, "starter" if i == 0 else f"secondary:{i - 1}", s.n_blocks, meta=True,
                        first_mlp_only=s.first_mlp_only, last_attn_only=s.last_attn_only)
        m.load_weights(chunks["starter"] if i == 0 else chunks["secondary"][i - 1])
        m.set_kv_cache(1)
        mods.append(m.eval())
    idx = torch.tensor([[1, 5, 9, 2]])
    full.set_kv_cache(1)
    with torch.no_grad():
        pos = torch.arange(4)
        ref = full(idx, pos)
        x = mods[0](idx, pos)
        for m in mods[1:]:
            x = m(x, pos)
        out = mods[0].head(x)
        torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-5)
        nxt = torch.tensor([[7]])
        pos = torch.tensor([4])
        ref = full(nxt, pos)
        x = mods[0](nxt, pos)
        for m in mods[1:]:
            x = m(x, pos)
        torch.testing.assert_close(mods[0].head(x), ref, rtol=1e-4, atol=1e-5)
