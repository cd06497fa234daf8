"""The sorting network of the two-operand CSG fast path (render.cuh csgPair) must visit crossings in the order of the
reference's stable sort.

Reference: four stable insertions ("after every element that is not greater", = F# Seq.sortBy over A's hits then B's,
Csg.fs:74-94; what evalCsg does).  Kernel: the four slots (absent ones = +inf, marked invalid) through an odd-even
transposition network whose exchanges swap neighbours only when the later key is strictly smaller.  This restates both
in Python and compares the sequence of valid ids on random inputs with many ties, zeros, infinities and absent slots.
(NaN keys are excluded: both builds document NaN ordering as outside the parity bar)"""
import itertools
import math
import random


def by_insertion(slots):
    mt, mid = [], []
    for t, ident in slots:
        if ident is None:
            continue
        p = sum(1 for x in mt if not (t < x))
        mt.insert(p, t)
        mid.insert(p, ident)
    return mid


def by_network(slots):
    mt = [t if ident is not None else math.inf for t, ident in slots]
    mid = [ident for _, ident in slots]
    for i, j in ((0, 1), (2, 3), (1, 2), (0, 1), (2, 3), (1, 2)):
        if mt[j] < mt[i]:
            mt[i], mt[j] = mt[j], mt[i]
            mid[i], mid[j] = mid[j], mid[i]
    return [x for x in mid if x is not None]


def test_exhaustive_small_keys():
    keys = [-1.0, 0.0, 0.5, 0.5, 2.0, math.inf]
    for ts in itertools.product(keys, repeat=4):
        for present in itertools.product((False, True), repeat=4):
            # a leaf reports hit 1 only after hit 0: (a0, a1, b0, b1) with a1 => a0, b1 => b0
            if (present[1] and not present[0]) or (present[3] and not present[2]):
                continue
            slots = [(ts[k], "ab"[k // 2] + str(k % 2) if present[k] else None) for k in range(4)]
            assert by_insertion(slots) == by_network(slots), slots


def test_random_with_ties():
    rnd = random.Random(3)
    for _ in range(20000):
        pool = [rnd.choice([0.0, 1.0, 1.0, 2.5, -3.0, math.inf]) if rnd.random() < 0.6 else rnd.uniform(-5, 5) for _ in range(4)]
        na, nb = rnd.choice([0, 1, 2]), rnd.choice([0, 1, 2])
        slots = [(pool[0], "a0" if na > 0 else None), (pool[1], "a1" if na > 1 else None),
                 (pool[2], "b0" if nb > 0 else None), (pool[3], "b1" if nb > 1 else None)]
        assert by_insertion(slots) == by_network(slots), slots


class _PushSink:
    """render.cuh PairSink: the last two crossings, newest first; csgPair reads them back in emission order."""

    def __init__(self):
        self.n, self.t0, self.t1, self.i0, self.i1 = 0, 0.0, 0.0, None, None

    def hit(self, t, ident):
        self.t1, self.i1 = self.t0, self.i0
        self.t0, self.i0 = t, ident
        self.n += 1

    def slots(self):
        first = (self.t1, self.i1) if self.n > 1 else (self.t0, self.i0)
        return [first if self.n > 0 else (0.0, None), (self.t0, self.i0) if self.n > 1 else (0.0, None)]


def test_push_front_sink_reads_back_in_emission_order():
    rnd = random.Random(5)
    for _ in range(2000):
        a, b = _PushSink(), _PushSink()
        na, nb = rnd.choice([0, 1, 2]), rnd.choice([0, 1, 2])
        ha = [(rnd.choice([0.0, 1.0, 1.0, -2.0]), "a%d" % k) for k in range(na)]
        hb = [(rnd.choice([0.0, 1.0, 1.0, -2.0]), "b%d" % k) for k in range(nb)]
        for t, i in ha:
            a.hit(t, i)
        for t, i in hb:
            b.hit(t, i)
        emitted = [(t, i) for t, i in ha] + [(0.0, None)] * (2 - na) + [(t, i) for t, i in hb] + [(0.0, None)] * (2 - nb)
        assert a.slots() + b.slots() == emitted
        assert by_network(a.slots() + b.slots()) == by_insertion(emitted)
