#!/usr/bin/env bash
# same-box A/B of decode-kernel tuning knobs (N=1 Llama-3-8B); prints tok/s per configuration
b() { python bench.py --steps 64 --warmup 8 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'])"; }
echo -n "prev lib:            "; MDI_OPS_LIB=$PWD/mdi_llm_b200/ops/build/ab_prev.so b
echo -n "current:             "; b
echo -n "gate_up grid 256:    "; MDI_CTAS_GATE_UP=-256 b
echo -n "gate_up grid 296:    "; MDI_CTAS_GATE_UP=-296 b
echo -n "gate_up grid 448:    "; MDI_CTAS_GATE_UP=-448 b
echo -n "qkv grid 296:        "; MDI_CTAS_QKV=-296 b
echo -n "down grid 296:       "; MDI_CTAS_DOWN=-296 b
echo -n "o_proj grid 148:     "; MDI_CTAS_O_PROJ=-148 b
echo -n "lm_head grid 334:    "; MDI_CTAS_LM_HEAD=-334 b
echo -n "current again:       "; b
