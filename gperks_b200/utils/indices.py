"""Index/set helpers used by history matching (behaviour of GPErks/utils/indices.py)."""
import numpy as np


def diff(l1, l2):
    return list(set(l1) - set(l2))


def inters(l1, l2):
    return list(set(l1) & set(l2))


def inters_many(L):
    return list(set.intersection(*map(set, L)))


def union_many(L):
    return list(set.union(*map(set, L)))


def _widest_axis(X):
    spans = X.max(axis=0) - X.min(axis=0)
    j = int(np.argmax(spans))
    return j, spans[j]


def part_and_select(P, N):
    """Greedy bisection of the point cloud ``P`` into ``N`` cells along each cell's widest axis; returns the
    cells and, per cell, the member closest to the cell centroid."""
    cells = [P]
    while len(cells) < N:
        widths = [_widest_axis(C)[1] for C in cells]
        if not any(widths):
            break
        k = int(np.argmax(widths))
        C = cells.pop(k)
        j, _ = _widest_axis(C)
        mid = 0.5 * (C[:, j].max() + C[:, j].min())
        cells.append(C[C[:, j] <= mid])
        cells.append(C[C[:, j] > mid])
    picks = [C[np.argmin(np.linalg.norm(C - C.mean(axis=0), axis=1))] for C in cells]
    return cells, np.stack(picks)


def whereq_whernot(X, SX):
    """Row indices of ``X`` matching the rows of ``SX`` (first match each) and the complementary indices."""
    hit = []
    for row in SX:
        match = np.where((X == row).all(axis=1))[0]
        hit.append(int(match[0]))
    rest = sorted(diff(range(X.shape[0]), hit))
    return hit, rest


def filter_zscore(X, thre):
    z = np.abs((X - X.mean(axis=0)) / X.std(axis=0))
    out = sorted(set(np.where(z > thre)[0].tolist()))
    keep = diff(range(X.shape[0]), out)
    return keep, out
