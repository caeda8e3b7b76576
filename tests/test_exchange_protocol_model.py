"""A discrete-event model of the flag protocol of csrc/nts_exchange.cu, run under random
interleavings: P ranks, each with a main stream and a side stream that execute their operations in order but at
arbitrary relative speeds; windows of n_buffers epoch-alternating buffers; pushed[] / consumed[] epoch flags.

It checks the two properties the design argues for:
  * safety  - whenever a rank aggregates from its window at epoch e, every slot holds the data of epoch e (no pusher has
              overwritten a buffer that is still being read, and no reader runs ahead of its pushers);
  * liveness - every schedule terminates (waits are only ever on strictly earlier epochs, so there is no cycle),
for 1 and 2 epoch buffers, 2-4 ranks, pipeline / merged / per-rank mixed receive strategies (a backward call runs
the same protocol with the roles of the two staging layouts swapped, so one call type is modelled).
This is a model of the protocol, not of the CUDA code; the CUDA path is covered by tests/test_multi_gpu.py."""
import random

import pytest


class Rank:
    def __init__(self, p, P, n_buffers):
        self.p, self.P, self.nb = p, P, n_buffers
        self.pushed = [0] * P        # pushed[j]: last epoch whose rows from j have landed in my window
        self.consumed = [0] * P      # consumed[j]: last epoch whose window contents j has finished reading
        self.window = [[0] * P for _ in range(n_buffers)]   # window[buf][j] = epoch of the data j stored there
        self.main, self.side = [], []                       # operation queues (closures returning True when done)


def build(P, n_buffers, epochs, merged, rng):
    ranks = [Rank(p, P, n_buffers) for p in range(P)]
    errors = []
    for e in range(1, epochs + 1):
        buf = e % n_buffers
        wait_epoch = e - n_buffers if e > n_buffers else 0
        for r in ranks:
            p = r.p
            # ---- side stream: push to every peer in ring order p-1, p-2, ...
            for s in range(1, P):
                j = (p - s) % P

                def push(j=j, p=p, e=e, buf=buf, wait_epoch=wait_epoch, r=r):
                    if r.consumed[j] < wait_epoch:      # peer j has not finished with the buffer I am about to overwrite
                        return False
                    ranks[j].window[buf][p] = e         # the rows land ...
                    ranks[j].pushed[p] = e              # ... then the flag (release)
                    return True
                r.side.append(push)
            # ---- main stream: read the window (per partition or all at once), then tell everybody
            order = [(p + s) % P for s in range(1, P)]
            groups = [order] if merged(e, p) else [[i] for i in order]
            for g in groups:
                def read(g=g, e=e, buf=buf, r=r):
                    if any(r.pushed[i] < e for i in g):
                        return False
                    for i in g:
                        if r.window[buf][i] != e:
                            errors.append("rank %d epoch %d: slot of %d holds epoch %d" % (r.p, e, i, r.window[buf][i]))
                    return True
                r.main.append(read)

            def done(p=p, e=e):
                for j in range(P):
                    if j != p:
                        ranks[j].consumed[p] = e
                return True
            r.main.append(done)

            # the main stream does not return to the caller before the side stream has sent this epoch (ev_comm)
            def join(r=r, target=len(r.side)):
                return r.side_done >= target
            r.main.append(join)
    for r in ranks:
        r.side_done = 0
    return ranks, errors


def run(ranks, rng, max_steps=200000):
    heads = {(r.p, k): 0 for r in ranks for k in ("main", "side")}
    queues = {(r.p, "main"): r.main for r in ranks}
    queues.update({(r.p, "side"): r.side for r in ranks})
    by_p = {r.p: r for r in ranks}
    for _ in range(max_steps):
        live = [k for k in heads if heads[k] < len(queues[k])]
        if not live:
            return True
        rng.shuffle(live)
        # a random subset of the streams gets to try its next operation; if none of THEM can move, everybody tries
        # once - only if nobody at all can move is it a deadlock
        progressed = False
        for attempt in (live[: max(1, len(live) // 2)], live):
            for k in attempt:
                if queues[k][heads[k]]():
                    heads[k] += 1
                    if k[1] == "side":
                        by_p[k[0]].side_done += 1
                    progressed = True
            if progressed:
                break
        if not progressed:
            return False
    return False


@pytest.mark.parametrize("P", [2, 3, 4])
@pytest.mark.parametrize("n_buffers", [1, 2])
@pytest.mark.parametrize("strategy", ["pipeline", "merged", "mixed"])
def test_flag_protocol_is_safe_and_live(P, n_buffers, strategy):
    for seed in range(25):
        rng = random.Random(1000 * P + 10 * n_buffers + seed)
        merged = {"pipeline": lambda e, p: False, "merged": lambda e, p: True,
                  "mixed": lambda e, p: (e + p) % 2 == 0}[strategy]      # every rank decides for itself
        ranks, errors = build(P, n_buffers, 7, merged, rng)
        assert run(ranks, rng), "deadlock: P=%d n_buffers=%d %s seed %d" % (P, n_buffers, strategy, seed)
        assert not errors, errors[:3]


def test_the_model_detects_a_missing_consumed_wait():
    """Sanity of the model itself: without the consumed wait a fast pusher clobbers a buffer that is still unread."""
    hit = False
    for seed in range(200):
        rng = random.Random(seed)
        ranks, errors = build(3, 1, 6, lambda e, p: False, rng)
        for r in ranks:                        # drop the wait: pretend everything has always been consumed
            r.consumed = [10 ** 6] * 3
        for r in ranks:                        # (done() would overwrite it: make it a no-op)
            r.main = [op for op in r.main if op.__name__ != "done"]
        run(ranks, rng)
        hit = hit or bool(errors)
    assert hit
