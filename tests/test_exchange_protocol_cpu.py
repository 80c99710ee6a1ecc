"""Model check of the peer-exchange protocol of drsa_finish_step_p2p (csrc/retract_fused.cu: peer_exchange) on the CPU.

Every rank repeats: read its exchange counter s -> push its share into slot [s & 1][rank] of every peer -> signal every peer's
arrival counter of parity s & 1 -> wait until its own counter of that parity shows world - 1 arrivals -> read all slots of that
parity -> rewind that counter and advance s.  The claim in DESIGN.md section 5 is that TWO parities suffice, i.e. that under
any interleaving of the ranks' actions (a) a slot is never overwritten before its owner has read it and (b) what a rank reads
is the peers' data of the same exchange.  The model executes the ranks' atomic actions in random order (thousands of
schedules, including ones where a rank runs as far ahead as the protocol lets it) and checks both.  This is the host-side
counterpart of the 2-GPU test in tests/test_gpu_dist.py, which can only see the schedules the hardware happens to produce."""
import random

import pytest


class Rank:
    def __init__(self, r, world):
        self.r, self.world = r, world
        self.step = 0                      # header[0]: exchange counter
        self.arrivals = [0, 0]             # header[16], header[32]
        self.inbox = [[None] * world for _ in range(2)]
        self.pc = 0                        # program counter inside one exchange
        self.push_to = 0
        self.done = 0

    def data(self, step):
        return (self.r, step)


def run_schedule(world, exchanges, rng, greedy_rank=None):
    ranks = [Rank(r, world) for r in range(world)]
    while any(rk.done < exchanges for rk in ranks):
        runnable = [rk for rk in ranks if rk.done < exchanges and
                    not (rk.pc == 2 and rk.arrivals[rk.step & 1] < world - 1)]
        assert runnable, "deadlock"
        if greedy_rank is not None and ranks[greedy_rank] in runnable and rng.random() < 0.9:
            rk = ranks[greedy_rank]        # one rank runs ahead whenever it can
        else:
            rk = rng.choice(runnable)
        par = rk.step & 1
        if rk.pc == 0:                     # push to the next peer (one store per action: pushes of different ranks interleave)
            peer = rk.push_to
            if peer != rk.r:
                slot = ranks[peer].inbox[par]
                # (a) the slot must not hold data its owner has not consumed yet
                assert slot[rk.r] is None, f"rank {rk.r} overwrites unread data of exchange {slot[rk.r][1]} at rank {peer}"
                slot[rk.r] = rk.data(rk.step)
            rk.push_to += 1
            if rk.push_to == world:
                rk.push_to, rk.pc = 0, 1
        elif rk.pc == 1:                   # one arrival per peer (the last CTA of the grid signals)
            for peer in range(world):
                if peer != rk.r:
                    ranks[peer].arrivals[par] += 1
                    assert ranks[peer].arrivals[par] <= world - 1, "more arrivals than peers on one parity"
            rk.pc = 2
        elif rk.pc == 2:                   # wait satisfied (checked above): read, rewind, advance
            for peer in range(world):
                if peer != rk.r:
                    got = rk.inbox[par][peer]
                    # (b) the data of the same exchange
                    assert got == (peer, rk.step), f"rank {rk.r} exchange {rk.step} read {got} from rank {peer}"
                    rk.inbox[par][peer] = None
            rk.arrivals[par] = 0
            rk.step += 1
            rk.done += 1
            rk.pc = 0
    return ranks


@pytest.mark.parametrize("world", [2, 3, 8])
def test_two_parities_suffice_under_random_schedules(world):
    rng = random.Random(1234 + world)
    for trial in range(300 if world < 8 else 60):
        greedy = rng.randrange(world) if trial % 2 else None
        ranks = run_schedule(world, exchanges=12, rng=rng, greedy_rank=greedy)
        assert all(rk.step == 12 and rk.arrivals == [0, 0] for rk in ranks)


def test_model_detects_a_single_parity_protocol():
    """Sanity of the checker itself: with ONE parity a fast rank overwrites data its peer has not read."""
    rng = random.Random(7)
    failures = 0
    for trial in range(200):
        ranks = [Rank(r, 2) for r in range(2)]
        try:
            # same loop as run_schedule but every exchange uses parity 0
            while any(rk.done < 6 for rk in ranks):
                runnable = [rk for rk in ranks if rk.done < 6 and not (rk.pc == 2 and rk.arrivals[0] < 1)]
                rk = ranks[0] if (ranks[0] in runnable and rng.random() < 0.9) else rng.choice(runnable)
                if rk.pc == 0:
                    peer = 1 - rk.r
                    assert ranks[peer].inbox[0][rk.r] is None
                    ranks[peer].inbox[0][rk.r] = rk.data(rk.step)
                    rk.pc = 1
                elif rk.pc == 1:
                    ranks[1 - rk.r].arrivals[0] += 1
                    rk.pc = 2
                else:
                    got = rk.inbox[0][1 - rk.r]
                    assert got == (1 - rk.r, rk.step)
                    rk.inbox[0][1 - rk.r] = None
                    rk.arrivals[0] -= 1
                    rk.step += 1; rk.done += 1; rk.pc = 0
        except AssertionError:
            failures += 1
    assert failures > 0
