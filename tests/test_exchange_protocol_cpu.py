"""Model check of the peer-exchange protocol of drsa_finish_step_p2p (csrc/retract_fused.cu: peer_push / peer_read) on the CPU.

Every rank repeats: read its exchange counter s -> store its words {value, flag = s + 1} into slot [s & 1][rank] of every
peer, one word at a time, in any order relative to the other ranks' stores -> for every peer word poll slot [s & 1][peer]
until its flag equals s + 1 and take the value -> advance s.  There is no other signal.  The claim in DESIGN.md section 5 is
that TWO parities suffice, i.e. that under any interleaving of the ranks' actions (a) a word is never overwritten before its
owner has read it and (b) what a rank reads is the peers' data of the same exchange.  The model executes the ranks' atomic
actions (single word stores, single word polls) in random order -- thousands of schedules, including ones where a rank runs as
far ahead as the protocol lets it -- and checks both.  This is the host-side counterpart of the 2-GPU test in
tests/test_gpu_dist.py, which can only see the schedules the hardware happens to produce."""
import random

import pytest

WORDS = 3                                  # words per share: enough to interleave partial pushes


class Rank:
    def __init__(self, r, world, parities=2):
        self.r, self.world, self.parities = r, world, parities
        self.step = 0                      # header[0]: exchange counter
        # inbox[parity][source][word] = (value, flag, consumed)
        self.inbox = [[[(None, 0, True)] * WORDS for _ in range(world)] for _ in range(parities)]
        self.to_push = []                  # (peer, word) stores still to make in this exchange
        self.to_read = []                  # (peer, word) words still to read in this exchange
        self.started = False
        self.done = 0

    def begin(self, rng):
        self.to_push = [(p, w) for p in range(self.world) if p != self.r for w in range(WORDS)]
        self.to_read = list(self.to_push)
        rng.shuffle(self.to_push)          # CTAs of the grid store in no particular order
        rng.shuffle(self.to_read)
        self.started = True


def act(rk, ranks, rng):
    """One atomic action of rank rk; returns False when it could only poll without success."""
    if not rk.started:
        rk.begin(rng)
    par, flag = rk.step % rk.parities, rk.step + 1
    # the kernel's threads push and then read, but different CTAs are at different points: pick either kind of action
    if rk.to_push and (not rk.to_read or rng.random() < 0.6):
        peer, w = rk.to_push.pop()
        old = ranks[peer].inbox[par][rk.r][w]
        # (a) the word must not hold data its owner has not consumed yet
        assert old[2], f"rank {rk.r} overwrites the unread word of exchange {old[1] - 1} at rank {peer}"
        ranks[peer].inbox[par][rk.r][w] = ((rk.r, rk.step, w), flag, False)
        return True
    progressed = False
    for i, (peer, w) in enumerate(rk.to_read):
        val, f, _ = rk.inbox[par][peer][w]
        if f == flag:
            # (b) the data of the same exchange
            assert val == (peer, rk.step, w), f"rank {rk.r} exchange {rk.step} read {val} from rank {peer}"
            rk.inbox[par][peer][w] = (val, f, True)
            rk.to_read.pop(i)
            progressed = True
            break
    if not rk.to_push and not rk.to_read:  # kernel end: every word pushed and read -> advance the counter
        rk.step += 1
        rk.done += 1
        rk.started = False
        return True
    return progressed


def run_schedule(world, exchanges, rng, greedy_rank=None, parities=2):
    ranks = [Rank(r, world, parities) for r in range(world)]
    idle = 0
    while any(rk.done < exchanges for rk in ranks):
        live = [rk for rk in ranks if rk.done < exchanges]
        if greedy_rank is not None and ranks[greedy_rank] in live and rng.random() < 0.9:
            rk = ranks[greedy_rank]        # one rank runs ahead whenever it can
        else:
            rk = rng.choice(live)
        idle = 0 if act(rk, ranks, rng) else idle + 1
        assert idle < 10000, "deadlock"
    return ranks


@pytest.mark.parametrize("world", [2, 3, 8])
def test_two_parities_suffice_under_random_schedules(world):
    rng = random.Random(1234 + world)
    for trial in range(200 if world < 8 else 30):
        greedy = rng.randrange(world) if trial % 2 else None
        ranks = run_schedule(world, exchanges=10, rng=rng, greedy_rank=greedy)
        assert all(rk.step == 10 for rk in ranks)
        assert all(w[2] for rk in ranks for par in rk.inbox for src in par for w in src)      # nothing left unread


def test_model_detects_a_single_parity_protocol():
    """Sanity of the checker itself: with ONE parity a fast rank overwrites words its peer has not read."""
    rng = random.Random(7)
    failures = 0
    for trial in range(200):
        try:
            run_schedule(2, exchanges=6, rng=rng, greedy_rank=0, parities=1)
        except AssertionError:
            failures += 1
    assert failures > 0
