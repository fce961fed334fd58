"""Error containment for the long-running node loops.

Parity: reference ``src/sub/utils/context_managers.py:16-56`` (``catch_loop_errors``): on any
exception inside the loop, clear the ``running`` event, set/clear the listed events (e.g. stop
the spinner thread) and re-raise; ``KeyboardInterrupt`` is swallowed after the same cleanup.
"""
from __future__ import annotations

import threading
from contextlib import contextmanager
from typing import Iterable, Iterator, Optional

__all__ = ["catch_loop_errors"]


@contextmanager
def catch_loop_errors(
    running_event: threading.Event,
    event_to_be_set: Optional[Iterable[threading.Event]] = None,
    event_to_be_cleared: Optional[Iterable[threading.Event]] = None,
) -> Iterator[None]:
    def cleanup() -> None:
        running_event.clear()
        for ev in event_to_be_set or ():
            ev.set()
        for ev in event_to_be_cleared or ():
            ev.clear()

    try:
        yield
    except KeyboardInterrupt:
        cleanup()
        print("Node was stopped!")
    except Exception:
        cleanup()
        raise
    else:
        for ev in event_to_be_set or ():
            ev.set()
