"""Adversarial inputs for libstdc++'s std::nth_element (McIlroy's "antiquicksort" adversary played against a
Python restatement of introselect): value sequences on which the median-of-three Hoare partitions keep
discarding O(1) elements until the depth limit 2*lg(n) is exhausted and std::__heap_select takes over.
The kd-tree builders reach that code through the median rule (nth at the middle) — the device build's
sequential fallback (seq_heap_select in build.cu) and the oracle's restatement are only exercised by such data.

    python oracle/make_killers.py   ->  tests/data/introselect_killers.npz

TEST INFRASTRUCTURE ONLY.
"""
import os

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "data",
                   "introselect_killers.npz")


class HeapSelectReached(Exception):
    pass


def unguarded_partition(a, first, last, pivot_pos):
    pv = a[pivot_pos]
    while True:
        while a[first] < pv:
            first += 1
        last -= 1
        while pv < a[last]:
            last -= 1
        if not first < last:
            return first
        a[first], a[last] = a[last], a[first]
        first += 1


def move_median_to_first(a, result, x, y, z):
    if a[x] < a[y]:
        pick = y if a[y] < a[z] else (z if a[x] < a[z] else x)
    elif a[x] < a[z]:
        pick = x
    elif a[y] < a[z]:
        pick = z
    else:
        pick = y
    a[result], a[pick] = a[pick], a[result]


def introselect(a, first, nth, last):
    """std::__introselect up to the point where it would call __heap_select."""
    depth = 2 * ((last - first).bit_length() - 1)
    while last - first > 3:
        if depth == 0:
            raise HeapSelectReached
        depth -= 1
        mid = first + (last - first) // 2
        move_median_to_first(a, first, first + 1, mid, last - 1)
        cut = unguarded_partition(a, first + 1, last, first)
        if cut <= nth:
            first = cut
        else:
            last = cut


def killer(n, nth):
    gas = n
    val = [gas] * n
    st = {"solid": 0, "cand": 0}

    class Key:
        __slots__ = ("i",)

        def __init__(self, i):
            self.i = i

        def __lt__(self, o):
            x, y = self.i, o.i
            if val[x] == gas and val[y] == gas:
                z = x if x == st["cand"] else y
                val[z] = st["solid"]
                st["solid"] += 1
            if val[x] == gas:
                st["cand"] = x
            elif val[y] == gas:
                st["cand"] = y
            return val[x] < val[y]

    try:
        introselect([Key(i) for i in range(n)], 0, nth, n)
    except HeapSelectReached:
        pass
    return np.array([v if v != gas else n - 1 for v in val], dtype=np.float32)


def reaches_heap_select(vals, nth):
    class P:
        __slots__ = ("v",)

        def __init__(self, v):
            self.v = v

        def __lt__(self, o):
            return self.v < o.v
    try:
        introselect([P(v) for v in vals], 0, nth, len(vals))
    except HeapSelectReached:
        return True
    return False


def main():
    out = {}
    for n in (300, 1000, 5000, 40000):
        v = killer(n, n // 2)
        assert reaches_heap_select(v.tolist(), n // 2), n
        out["n%d" % n] = v
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
