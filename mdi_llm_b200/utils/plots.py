"""Tokens-vs-time output.

Parity: reference ``src/sub/utils/plots.py:12-51`` (``plot_tokens_per_time``).  matplotlib is an
optional dependency here (absent on the GPU image): without it the points are written as CSV next
to the requested image path and a text sparkline is printed, so ``-p`` never fails.
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import Optional, Sequence, Tuple, Union

__all__ = ["plot_tokens_per_time", "have_matplotlib", "write_points_csv"]


def have_matplotlib() -> bool:
    try:
        import matplotlib  # noqa: F401

        return True
    except Exception:  # noqa: BLE001
        return False


def write_points_csv(tok_time: Sequence[Tuple[int, float]], path: Union[str, Path]) -> Path:
    """``time,n_tokens`` rows — column order of the reference's CSV (starter.py:78-82)."""
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    with open(path, "w") as f:
        for n, t in tok_time:
            f.write(f"{t},{n}\n")
    return path


def plot_tokens_per_time(tok_time: Sequence[Tuple[int, float]], out_path: Optional[Union[str, Path]] = None,
                         disp: bool = False, label: Optional[str] = None) -> Optional[Path]:
    if not tok_time:
        return None
    if have_matplotlib():
        import matplotlib

        if not disp:
            matplotlib.use("Agg")
        import matplotlib.pyplot as plt

        fig = plt.figure(figsize=(12, 6))
        plt.plot([t for _, t in tok_time], [n for n, _ in tok_time], label=label)
        plt.xlabel("Time (s)")
        plt.ylabel("Tokens")
        plt.title("Number of generated tokens vs. time")
        plt.grid()
        if label:
            plt.legend()
        plt.tight_layout()
        saved = None
        if out_path is not None:
            os.makedirs(os.path.dirname(str(out_path)) or ".", exist_ok=True)
            fig.savefig(out_path)
            saved = Path(out_path)
        if disp:
            plt.show()
        plt.close(fig)
        return saved
    n_last, t_last = tok_time[-1]
    rate = n_last / t_last if t_last > 0 else float("nan")
    print(f"[plots] matplotlib not installed: {n_last} tokens in {t_last:.3f} s ({rate:.1f} tok/s)")
    if out_path is not None:
        return write_points_csv(tok_time, Path(out_path).with_suffix(".csv"))
    return None
