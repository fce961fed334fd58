"""Reference arm of bench.py (filled in by baseline/ harness)."""
from __future__ import annotations

from typing import Any, Dict


def run_reference(args: Any) -> Dict[str, Any]:
    return {"impl": "reference", "unavailable": "reference harness not installed yet"}
