#!/usr/bin/env bash
# same-box A/B of decode-kernel changes (N=1 Llama-3-8B); prints tok/s per configuration
b() { python bench.py --steps 64 --warmup 8 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'])"; }
AB=$PWD/mdi_llm_b200/ops/build
echo -n "prev lib:                 "; MDI_OPS_LIB=$AB/ab_prev.so b
echo -n "current (new attn):       "; b
echo -n "current, old attention:   "; MDI_OPS_LIB=$AB/ab_oldattn.so b
for mb in 8 16 32 64; do echo -n "current + L2 prefetch $mb MB: "; MDI_L2_PF_MB=$mb b; done
echo -n "old attn + L2 pf 32:      "; MDI_L2_PF_MB=32 MDI_OPS_LIB=$AB/ab_oldattn.so b
echo -n "current again:            "; b
