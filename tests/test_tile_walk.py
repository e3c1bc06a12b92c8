"""CPU restatement of the tile kernels' walk over the active-tile list (phases_tiled.cuh: TileWalk, walk_begin /
walk_prefetch / walk_next): the first `fixed8` eighths of the list, rounded to whole rounds of the resident warps, are
dealt with a fixed stride, the rest is handed out by a ticket counter, the next ticket drawn at the start of a tile.
Whatever the order in which the warps finish their tiles, every list entry has to be processed exactly once.  (The
CUDA code itself is exercised by tests/test_gpu_parity.py; this pins the index arithmetic.)"""
import numpy as np
import pytest


class Walk:
    def __init__(self, first, stride, n_act, fixed8, counter):
        self.stride, self.n_act, self.counter = stride, n_act, counter
        self.ticketed = counter is not None
        self.n_fixed = ((n_act // stride) * fixed8 // 8) * stride if self.ticketed else n_act
        self.a, self.raw = first, 0
        if self.ticketed and self.a >= self.n_fixed:
            self.a = self.n_fixed + self.draw()

    def draw(self):
        t = self.counter[0]
        self.counter[0] += 1
        return t

    def prefetch(self):                       # at the top of the loop body
        if self.ticketed and self.a >= self.n_fixed:
            self.raw = self.draw()

    def next(self):
        if not self.ticketed:
            self.a += self.stride
        elif self.a >= self.n_fixed:
            self.a = self.n_fixed + self.raw
        else:
            self.a += self.stride
            if self.a >= self.n_fixed:
                self.a = self.n_fixed + self.draw()


@pytest.mark.parametrize("n_act", [0, 1, 5, 37, 1000, 4099])
@pytest.mark.parametrize("fixed8", [0, 4, 7, 8])
@pytest.mark.parametrize("ticketed", [False, True])
def test_every_list_entry_is_processed_exactly_once(n_act, fixed8, ticketed):
    rng = np.random.default_rng(n_act * 31 + fixed8)
    n_warps = 24
    counter = [0] if ticketed else None
    walks = [Walk(w, n_warps, n_act, fixed8, counter) for w in range(n_warps)]
    seen = np.zeros(n_act, dtype=int)
    live = [w for w in walks if w.a < n_act]
    in_body = {}
    while live:
        w = live[rng.integers(len(live))]       # warps finish their tiles in any order
        if id(w) not in in_body:                # top of the loop body: the tile is taken, the next ticket drawn
            w.prefetch()
            seen[w.a] += 1
            in_body[id(w)] = True
        else:                                   # the tile is done
            del in_body[id(w)]
            w.next()
            if w.a >= n_act:
                live.remove(w)
    assert (seen == 1).all()
    if ticketed:
        n_fixed = ((n_act // n_warps) * fixed8 // 8) * n_warps
        assert n_fixed % n_warps == 0 and n_fixed <= n_act
        assert counter[0] >= n_act - n_fixed    # every entry behind the fixed part went through a ticket
