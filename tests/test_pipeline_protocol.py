"""Discrete-event model of the accumulator hand-off in score_select_kernel (host-only).

Three TMEM accumulator stages, mbarrier parity waits, MMA issuer warp(s) -> tensor pipe -> epilogue warp group(s).
An mbarrier parity wait is only sound for a waiter that can never fall two phases behind the barrier; the model runs
the protocol under random interleavings and checks that every tile is read exactly once, after its MMAs completed
and before its stage is overwritten.  It documents why the kernel uses ONE in-order issuer as soon as the epilogue
warps are split into groups (evavos_b200/csrc/score_tc.cu), and that the shipped configurations are safe.
"""
import random

import pytest


class Barrier:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def passed(self, parity):          # mbarrier.try_wait.parity: "the phase with this parity has completed"
        return (self.phase & 1) != parity


def shared_stages(i):
    """The kernel's mapping: three stages used round-robin by all tiles -> (stage, use count of that stage)."""
    return i % 3, i // 3


def per_group_stages(i):
    """Two independent pipelines (a round-2 candidate: needs 4 x 128 TMEM columns, i.e. the query operand in
    shared memory): tiles of parity g use stages 2g and 2g + 1 alternately."""
    g, j = i % 2, i // 2
    return 2 * g + j % 2, j // 2


def simulate(n_iter, n_issuers, groups, seed, steps=200000, stages=shared_stages, n_stages=3):
    """groups: list of functions i -> bool (does this epilogue group visit iteration i).  Returns None or an error."""
    rng = random.Random(seed)
    full = [Barrier(1) for _ in range(n_stages)]
    empty = [Barrier(1) for _ in range(n_stages)]   # one arrival per visiting group (its 8 warps move together)
    content = [None] * n_stages                 # iteration whose scores the stage holds (or is being written with)
    prev_user = {}                              # iteration -> the iteration that used its stage before it
    last = {}
    for i in range(n_iter):
        st, _ = stages(i)
        prev_user[i] = last.get(st)
        last[st] = i
    ready = [False] * n_iter                    # MMAs of iteration i completed
    consumed = [False] * n_iter
    issue_pos = [w for w in range(n_issuers)]   # next iteration of issuer w (w, w + n_issuers, ...)
    commits = [[] for _ in range(n_issuers)]    # per issuer: iterations whose commit has not fired yet (in order)
    visit_lists = [[i for i in range(n_iter) if g(i)] for g in groups]
    visit_pos = [0] * len(groups)

    for _ in range(steps):
        moves = []
        for w in range(n_issuers):
            i = issue_pos[w]
            if i < n_iter and empty[stages(i)[0]].passed((stages(i)[1] & 1) ^ 1):
                moves.append(("issue", w))
            if commits[w]:
                moves.append(("complete", w))
        for g, lst in enumerate(visit_lists):
            if visit_pos[g] < len(lst):
                i = lst[visit_pos[g]]
                if full[stages(i)[0]].passed(stages(i)[1] & 1):
                    moves.append(("visit", g))
        if not moves:
            done = all(consumed) and all(p >= n_iter for p in issue_pos)
            return None if done else "deadlock"
        kind, who = rng.choice(moves)
        if kind == "issue":
            i = issue_pos[who]
            if prev_user[i] is not None and not consumed[prev_user[i]]:
                return f"stage {stages(i)[0]} overwritten by tile {i} before tile {prev_user[i]} was read"
            content[stages(i)[0]] = i
            commits[who].append(i)
            issue_pos[who] += n_issuers
        elif kind == "complete":               # the tensor pipe finishes the oldest tile of one issuing thread
            i = commits[who].pop(0)
            ready[i] = True
            full[stages(i)[0]].arrive()
        else:
            i = visit_lists[who][visit_pos[who]]
            if not ready[i] or content[stages(i)[0]] != i:
                return f"group {who} read stage {stages(i)[0]} for tile {i} on a stale parity"
            consumed[i] = True
            empty[stages(i)[0]].arrive()
            visit_pos[who] += 1
    return "did not finish"


LOCKSTEP = [lambda i: True]
TWO_GROUPS = [lambda i: i % 2 == 0, lambda i: i % 2 == 1]
THREE_GROUPS = [lambda i: i % 3 == 0, lambda i: i % 3 == 1, lambda i: i % 3 == 2]


@pytest.mark.parametrize("name,n_issuers,groups", [("lock-step, two issuers (EVAVOS_GROUPS=1)", 2, LOCKSTEP),
                                                   ("two groups, one issuer (default)", 1, TWO_GROUPS),
                                                   ("three groups, one issuer (EVAVOS_GROUPS=3)", 1, THREE_GROUPS)])
def test_shipped_configurations_are_safe(name, n_issuers, groups):
    for seed in range(300):
        for n_iter in (1, 2, 5, 12, 46):
            assert simulate(n_iter, n_issuers, groups, seed) is None, (name, seed, n_iter)


def test_two_issuers_with_split_groups_is_unsafe():
    """The configuration that hung on the GPU: the model finds the stale-parity read / overwrite as well."""
    errors = {simulate(46, 2, TWO_GROUPS, seed) for seed in range(300)}
    assert any(e is not None for e in errors), errors


def test_two_independent_pipelines_would_be_safe():
    """Round-2 candidate: two issuers again, each feeding its own epilogue group through its own two stages."""
    for seed in range(300):
        for n_iter in (1, 2, 5, 12, 46):
            assert simulate(n_iter, 2, TWO_GROUPS, seed, stages=per_group_stages, n_stages=4) is None, (seed, n_iter)
