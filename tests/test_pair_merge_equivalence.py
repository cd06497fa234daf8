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
