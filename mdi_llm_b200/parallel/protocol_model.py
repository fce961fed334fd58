"""Explicit-state model of the device ring's flag protocol, small enough to enumerate EVERY interleaving.

What is modelled (the code it abstracts: ``ops/csrc/common.cuh`` ``hop_wait_one`` / ``hop_signal_copy``,
``ops/csrc/decode_misc.cu`` ``advance_step_kernel``, ``parallel/pipeline.py``):

* ``S`` stages in a ring, ``n`` sample slots, prefill (round 0) then ``R`` decode rounds.  Every stage owns one input
  row and one flag per slot; a stage's step for ``(slot, round)`` is: wait for its own ``flag[slot]``, read the row (twice:
  the first kernel stages it, a later kernel re-reads it as the residual), store the result into the NEXT stage's row
  (a plain, non-atomic store) and only then release the next stage's flag.  The stages run concurrently with no other
  synchronisation — exactly the situation of the GPUs of a box.
* flag values: the prefill message of a slot is published as ``1``, the message of decode round ``r`` as ``r + 1``;
  a secondary waits for ``round + 1``, the starter's head for ``round`` (the row that came back round the ring).
* abort: ``POISON`` is larger than every round number.  A stage that reads it, or whose watchdog expires, becomes
  *aborted* (sticky): it never waits again and publishes ``POISON`` instead of round numbers.  The host may overwrite a
  node's own flags with ``POISON`` (``DevicePipeline.poison`` — what ``PUT /stop`` and ``RingSession.abort`` do).

``explore`` walks the full interleaving graph (depth-first, memoised) and checks at every read that the row holds
exactly the message the step is about to consume (never a newer one: no overrun without back-pressure; never an older
one: no stale read), and at every terminal state that all live stages have finished (no deadlock).  After the FIRST
abort anywhere the generation is void — an aborted stage runs ahead without waiting and may overwrite rows that were not
consumed yet (counted as ``void_reads``); that is sound because every node's status word reaches the starter's session,
which raises ``RingError`` instead of returning tokens (``parallel/ring.py``).  Fault scenarios:
``spurious_trip`` (one watchdog fires although nothing is wrong), ``dead`` (a stage stops forever at an arbitrary point)
with recovery either by the successor's watchdog or by the host poisoning the live nodes.

The properties hold by construction — one message per slot is in flight round the ring, because the starter issues round
``r + 1`` of a slot only after round ``r`` came back — and this file is the executable form of that argument
(``tests/test_hop_protocol_model.py`` runs it; cross-GPU races in the kernels themselves are what compute-sanitizer's
racecheck / synccheck runs in ``profiles/`` cover)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

POISON = 0x7FFFFFFF

# micro-operations of one step
WAIT, READ1, READ2, WRITE, RELEASE = range(5)


@dataclass(frozen=True)
class Scenario:
    n_stages: int
    n_slots: int
    rounds: int                      # decode rounds after the prefill round
    spurious_trip: bool = False      # one watchdog expiry may happen at any blocked... or unblocked wait, anywhere
    dead: Optional[int] = None       # this stage may stop forever at an arbitrary point
    watchdog: bool = False           # blocked live stages may time out — only when nothing else can move (a long time)
    host_poison: bool = False        # after the death the host poisons every live node's own flags, each at any time
    release_before_write: bool = False  # a deliberately BROKEN producer (flag first, row second): the checker must object


def program(stage: int, sc: Scenario) -> List[Tuple[int, int, int, int]]:
    """Static schedule of a stage: ``(slot, round, wait_value or -1, signal_value)`` per step, prefill steps first."""
    steps = []
    for slot in range(sc.n_slots):  # prefill: the starter has nothing to wait for
        steps.append((slot, 0, -1 if stage == 0 else 1, 1))
    for r in range(1, sc.rounds + 1):
        for slot in range(sc.n_slots):
            steps.append((slot, r, r if stage == 0 else r + 1, r + 1))
    if stage == 0:  # the final head pass consumes the last message of every slot (no blocks, nothing sent)
        for slot in range(sc.n_slots):
            steps.append((slot, sc.rounds + 1, sc.rounds + 1, -1))
    return steps


class Violation(AssertionError):
    pass


def explore(sc: Scenario, max_states: int = 2_000_000) -> Dict[str, int]:
    """Enumerate every reachable state.  Raises :class:`Violation` on a bad read or a deadlock; returns counters."""
    S, n = sc.n_stages, sc.n_slots
    progs = [program(s, sc) for s in range(S)]
    # state: (pcs, micro-ops, aborted, flags, rows, dead_now, tripped, poisoned_mask, timeouts)
    flags0 = tuple(tuple(0 for _ in range(n)) for _ in range(S))
    rows0 = tuple(tuple(None for _ in range(n)) for _ in range(S))
    init = (tuple(0 for _ in range(S)), tuple(WAIT for _ in range(S)), tuple(False for _ in range(S)), flags0, rows0,
            False, False, 0, 0)
    seen = {init}
    stack = [init]
    stats = {"states": 0, "terminal": 0, "max_timeouts": 0, "aborted_terminals": 0, "void_reads": 0}

    def set2(t, i, j, v):
        row = list(t[i]); row[j] = v
        out = list(t); out[i] = tuple(row)
        return tuple(out)

    def set1(t, i, v):
        out = list(t); out[i] = v
        return tuple(out)

    while stack:
        st = stack.pop()
        stats["states"] += 1
        if stats["states"] > max_states:
            raise RuntimeError("state space larger than expected")
        pcs, ops, aborted, flags, rows, dead_now, tripped, poisoned, timeouts = st
        succ = []
        blocked = []
        live = [s for s in range(S) if not (dead_now and s == sc.dead)]
        for s in live:
            if pcs[s] >= len(progs[s]):
                continue
            slot, rnd, want, sig = progs[s][pcs[s]]
            op = ops[s]
            nxt = (s + 1) % S
            if op == WAIT:
                if want < 0 or aborted[s]:
                    succ.append((set1(pcs, s, pcs[s]), set1(ops, s, READ1), aborted, flags, rows, dead_now, tripped, poisoned, timeouts))
                elif flags[s][slot] >= want:
                    ab = set1(aborted, s, True) if flags[s][slot] == POISON else aborted
                    succ.append((pcs, set1(ops, s, READ1), ab, flags, rows, dead_now, tripped, poisoned, timeouts))
                else:
                    blocked.append(s)
                if sc.spurious_trip and not tripped and want >= 0 and not aborted[s]:  # the watchdog of this wait fires
                    succ.append((pcs, set1(ops, s, READ1), set1(aborted, s, True), flags, rows, dead_now, True, poisoned, timeouts))
            elif op in (READ1, READ2):
                if want >= 0 and not aborted[s]:
                    src = S - 1 if s == 0 else s - 1
                    expect = (slot, rnd - 1 if s == 0 else rnd, src)
                    if rows[s][slot] != expect:
                        # Once ANY stage has aborted the generation is void (an aborted stage stops waiting and runs ahead,
                        # so it may overwrite rows its successor has not consumed yet): every node's status word goes back
                        # to the starter's session, which raises RingError.  Before the first abort a wrong row is a bug.
                        if not any(aborted):
                            raise Violation(f"stage {s} step (slot {slot}, round {rnd}) read {rows[s][slot]} instead of {expect}")
                        stats["void_reads"] += 1
                succ.append((pcs, set1(ops, s, READ2 if op == READ1 else WRITE), aborted, flags, rows, dead_now, tripped, poisoned, timeouts))
            elif op in (WRITE, RELEASE):
                if sig < 0:  # final head pass: nothing to send
                    succ.append((set1(pcs, s, pcs[s] + 1), set1(ops, s, WAIT), aborted, flags, rows, dead_now, tripped, poisoned, timeouts))
                    continue
                do_row = (op == WRITE) != sc.release_before_write  # correct order: the row, then the flag
                f2 = flags if do_row else set2(flags, nxt, slot, POISON if aborted[s] else sig)
                r2 = set2(rows, nxt, slot, (slot, rnd, s)) if do_row else rows
                if op == WRITE:
                    succ.append((pcs, set1(ops, s, RELEASE), aborted, f2, r2, dead_now, tripped, poisoned, timeouts))
                else:
                    succ.append((set1(pcs, s, pcs[s] + 1), set1(ops, s, WAIT), aborted, f2, r2, dead_now, tripped, poisoned, timeouts))
        if sc.dead is not None and not dead_now and pcs[sc.dead] < len(progs[sc.dead]):
            succ.append((pcs, ops, aborted, flags, rows, True, tripped, poisoned, timeouts))  # the stage dies here
        if sc.host_poison and dead_now:
            for s in live:
                if not (poisoned >> s) & 1:
                    f2 = list(flags); f2[s] = tuple(POISON for _ in range(n))
                    succ.append((pcs, ops, aborted, tuple(f2), rows, dead_now, tripped, poisoned | (1 << s), timeouts))
        if not succ and sc.watchdog and blocked:  # nothing else can move: a watchdog expires
            for s in blocked:
                succ.append((pcs, set1(ops, s, READ1), set1(aborted, s, True), flags, rows, dead_now, tripped, poisoned, timeouts + 1))
        if not succ:
            unfinished = [s for s in live if pcs[s] < len(progs[s])]
            if unfinished:
                raise Violation(f"deadlock: stages {unfinished} cannot move (pcs {pcs}, flags {flags}, aborted {aborted})")
            stats["terminal"] += 1
            stats["max_timeouts"] = max(stats["max_timeouts"], timeouts)
            stats["aborted_terminals"] += int(any(aborted))
            continue
        for nx in succ:
            if nx not in seen:
                seen.add(nx)
                stack.append(nx)
    return stats
