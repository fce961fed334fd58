"""Exhaustive interleaving check of the device ring's flag protocol (mdi_llm_b200/parallel/protocol_model.py): message
integrity without back-pressure, no deadlock, and the abort / poison paths."""
import pytest

from mdi_llm_b200.parallel.protocol_model import Scenario, Violation, explore


@pytest.mark.parametrize("S,n,R", [(2, 1, 2), (2, 2, 2), (3, 2, 2), (3, 3, 1), (2, 3, 2), (4, 2, 1), (3, 1, 3)])
def test_every_interleaving_delivers_the_right_message(S, n, R):
    """n_samples <, = and > n_stages: a stage never reads a newer or an older row than the one its step consumes, and
    every stage finishes, although producers store into their successor's memory without asking."""
    st = explore(Scenario(S, n, R))
    assert st["terminal"] >= 1 and st["aborted_terminals"] == 0 and st["states"] > 20


def test_the_checker_catches_a_flag_released_before_the_row():
    with pytest.raises(Violation, match="read"):
        explore(Scenario(2, 2, 1, release_before_write=True))


@pytest.mark.parametrize("S,n,R", [(3, 2, 2), (2, 3, 1), (4, 1, 2)])
def test_one_spurious_watchdog_trip_drains_the_ring(S, n, R):
    """A single watchdog expiry anywhere (nothing else wrong): no other stage needs a timeout to finish — the poison
    published by the aborted stage satisfies every later wait round the ring — and no live read sees a wrong row."""
    st = explore(Scenario(S, n, R, spurious_trip=True))
    assert st["aborted_terminals"] > 0 and st["max_timeouts"] == 0


@pytest.mark.parametrize("dead", [0, 1, 2])
def test_dead_stage_recovered_by_the_watchdogs(dead):
    """A stage stops forever at an arbitrary point: every live stage still terminates, each with at most one expiry
    (aborted is sticky), and never consumes a wrong row before it aborts."""
    st = explore(Scenario(3, 2, 2, dead=dead, watchdog=True))
    assert st["terminal"] >= 1 and st["max_timeouts"] <= 2


@pytest.mark.parametrize("dead", [0, 2])
def test_dead_stage_recovered_by_host_poison_without_any_watchdog(dead):
    """`RingSession.abort` / `PUT /stop`: the host overwrites each live node's own flags with the poison value, at
    arbitrary moments relative to the producers' stores.  No watchdog is needed for the live stages to drain."""
    st = explore(Scenario(3, 2, 1, dead=dead, host_poison=True))
    assert st["terminal"] >= 1 and st["max_timeouts"] == 0
