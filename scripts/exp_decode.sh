#!/usr/bin/env bash
# same-box A/B of decode-path options (N=1 Llama-3-8B); prints tok/s, e2e tok/s and the watchdog status
b() { timeout 300 python bench.py --steps 64 --warmup 8 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], 'e2e', d['e2e']['value'], 'status', d.get('hop_watchdog_status'))"; }
echo -n "grid-completion deps (default): "; b
echo -n "flag deps (MDI_DEP_FLAGS=1):    "; MDI_DEP_FLAGS=1 b
echo -n "grid-completion deps again:     "; b
echo -n "flag deps again:                "; MDI_DEP_FLAGS=1 b
