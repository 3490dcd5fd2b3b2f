"""Pins for the kdtree2 restatement (oracle/letkf_oracle.cpp; module_kdtree2.f90).

The reference has no tests, so the oracle is pinned by
(1) the brute-force scan the reference itself ships as a self-check idea
    (kdtree2_r_nearest_brute_force, kd2:1755-1793);
(2) hand-derived closer-child-first DFS orders with truncation (kd2:1415-1428,1696-1706);
(3) an independent pure-Python restatement of build + search, written recursively from the
    Fortran text, compared on random small trees.
"""
import numpy as np
import pytest

from oracle import oracle as O


# ---------------------------------------------------------------- independent Python restatement
class _PyTree:
    def __init__(self, data, dim):
        self.d = np.asarray(data, np.float32)
        self.dim = dim
        self.n = len(self.d)
        self.ind = list(range(self.n))          # 0-based values, 0-based positions
        self.root = self._build(0, self.n - 1, None)

    def _spread(self, c, l, u):
        vals = [self.d[self.ind[i], c] for i in range(l, u + 1)]
        return np.float32(min(vals)), np.float32(max(vals))

    def _select(self, c, k, l, u):
        ind, d = self.ind, self.d
        while l < u:
            t, m = ind[l], l
            for i in range(l + 1, u + 1):
                if d[ind[i], c] < d[t, c]:
                    m += 1
                    ind[m], ind[i] = ind[i], ind[m]
            ind[l], ind[m] = ind[m], ind[l]
            if m <= k:
                l = m + 1
            if m >= k:
                u = m - 1

    def _build(self, l, u, parent):
        node = {"l": l, "u": u, "left": None, "right": None}
        if u - l <= 12:
            node["box"] = [self._spread(c, l, u) for c in range(self.dim)]
            node["cut_dim"] = -1
            return node
        box = []
        for c in range(self.dim):
            if parent is None or c == parent["cut_dim"]:
                box.append(self._spread(c, l, u))
            else:
                box.append(parent["box"][c])
        spreads = [np.float32(b[1] - b[0]) for b in box]
        c = int(np.argmax(spreads))             # first maximum, like maxloc
        m = (l + 1 + u + 1) // 2 - 1            # Fortran (l+u)/2 on 1-based bounds
        self._select(c, m, l, u)
        node["cut_dim"], node["box"] = c, box
        node["left"] = self._build(l, m, node)
        node["right"] = self._build(m + 1, u, node)
        node["cut_l"] = node["left"]["box"][c][1]
        node["cut_r"] = node["right"]["box"][c][0]
        node["cut_val"] = np.float32((node["cut_l"] + node["cut_r"]) / np.float32(2))
        node["box"] = [(min(a[0], b[0]), max(a[1], b[1]))
                       for a, b in zip(node["left"]["box"], node["right"]["box"])]
        return node

    def r_nearest(self, q, r2, nalloc):
        q = np.asarray(q, np.float32)
        r2 = np.float32(r2)
        out = []

        def leaf(node):
            for i in range(node["l"], node["u"] + 1):
                sd = np.float32(0)
                ok = True
                for c in range(self.dim):
                    df = np.float32(self.d[self.ind[i], c] - q[c])
                    sd = np.float32(sd + np.float32(df * df))
                    if sd > r2:
                        ok = False
                        break
                if not ok:
                    continue
                if len(out) >= nalloc:
                    return
                out.append((self.ind[i] + 1, sd))

        def bnd(x, lo, hi):
            if x > hi:
                return np.float32(np.float32(x - hi) ** 2)
            if x < lo:
                return np.float32(np.float32(lo - x) ** 2)
            return np.float32(0)

        def search(node):
            if node["left"] is None:
                leaf(node)
                return
            c = node["cut_dim"]
            if q[c] < node["cut_val"]:
                closer, farther = node["left"], node["right"]
                dis = np.float32(np.float32(node["cut_r"] - q[c]) ** 2)
            else:
                closer, farther = node["right"], node["left"]
                dis = np.float32(np.float32(node["cut_l"] - q[c]) ** 2)
            search(closer)
            if dis <= r2:
                for i in range(self.dim):
                    if i != c:
                        dis = np.float32(dis + bnd(q[i], *node["box"][i]))
                        if dis > r2:
                            return
                search(farther)

        search(self.root)
        return np.array([o[0] for o in out], np.int32), np.array([o[1] for o in out], np.float32)


def _pad3(a):
    out = np.zeros((len(a), 3), np.float32)
    out[:, :a.shape[1]] = a
    return out


# ---------------------------------------------------------------- (1) brute force
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("n", [1, 5, 13, 14, 27, 200, 5000])
def test_tree_equals_brute_force(dim, n):
    rng = np.random.default_rng(100 * dim + n)
    data = _pad3(rng.uniform(-10, 10, (n, dim)).astype(np.float32))
    if dim == 2:
        data[:, 2] = -1.0                       # what build_tree stores (loc:156)
    t = O.KdTree(data, dim)
    r2 = O.search_r2()
    for _ in range(25):
        q = rng.uniform(-11, 11, dim).astype(np.float32)
        idx, dis, tot = t.r_nearest(q, r2, n + 1)
        bidx, bdis = t.brute(q, r2)
        assert tot == len(bidx) == len(idx)
        o1, o2 = np.argsort(idx), np.argsort(bidx)
        assert np.array_equal(idx[o1], bidx[o2])
        assert np.array_equal(dis[o1].view(np.int32), bdis[o2].view(np.int32))  # bit-exact r2


def test_duplicates_and_boundary_inclusive():
    # sd <= r2 is inclusive (kd2:1680): a point exactly on the sphere is accepted
    data = _pad3(np.array([[3.0, 4.0]] * 20 + [[0.0, 0.0]] * 20 + [[6.0, 0.0]], np.float32))
    t = O.KdTree(data, 2)
    idx, dis, tot = t.r_nearest([0.0, 0.0], 25.0, 100)
    assert tot == 40 and set(idx) == set(range(1, 41))
    idx, _, tot = t.r_nearest([0.0, 0.0], np.nextafter(np.float32(25.0), np.float32(0)), 100)
    assert tot == 20 and set(idx) == set(range(21, 41))


# ---------------------------------------------------------------- (2) hand-derived DFS order
def test_truncation_keeps_first_hits_in_dfs_order_not_nearest():
    """26 collinear points x=1..26: root [1,26] splits at m=(1+26)/2=13 into two 13-point
    leaves (u-l = 12 <= bucket_size, kd2:737).  cut_val_left=13, cut_val_right=14,
    cut_val=13.5.  Query x=15.2, r2=9: closer child is the RIGHT leaf (qval >= cut_val,
    kd2:1420-1423) scanned in storage order -> 14,15,16,17,18; then the left leaf is visited
    because (13-15.2)^2 = 4.84 <= 9 (kd2:1433) -> 13."""
    data = _pad3(np.stack([np.arange(1, 27, dtype=np.float32), np.zeros(26, np.float32)], 1))
    t = O.KdTree(data, 2)
    d = t.dump()
    assert d["left"][0] >= 0 and d["cut_dim"][0] == 1
    assert d["cut_l"][0] == 13.0 and d["cut_r"][0] == 14.0 and d["cut_val"][0] == 13.5
    assert list(d["ind"]) == list(range(1, 27))
    q = [15.2, 0.0]
    idx, dis, tot = t.r_nearest(q, 9.0, 100)
    assert list(idx) == [14, 15, 16, 17, 18, 13] and tot == 6
    idx, _, tot = t.r_nearest(q, 9.0, 5)
    assert list(idx) == [14, 15, 16, 17, 18]      # 13 (distance 2.2) lost, 18 (2.8) kept
    assert tot >= 6                                # the walk continues after overflow (kd2:1697-1702)
    idx, _, _ = t.r_nearest(q, 9.0, 3)
    assert list(idx) == [14, 15, 16]              # nearest three would be 15, 16, 14
    # a query left of the cut visits the left leaf first
    idx, _, _ = t.r_nearest([12.9, 0.0], 4.0, 100)
    assert list(idx) == [11, 12, 13, 14]


def test_far_child_pruned_by_box_distance():
    """Second pruning test (kd2:1440-1448): the cut-plane distance passes but the node box in
    the other dimension is too far."""
    xs = np.arange(1, 27, dtype=np.float32)
    data = _pad3(np.stack([xs, np.zeros(26, np.float32)], 1))
    t = O.KdTree(data, 2)
    # y = 2.9: plane distance (13-13.6)^2 = .36 <= 9 but 0.36 + 2.9^2 = 8.77 <= 9 -> visited
    idx, _, _ = t.r_nearest([13.6, 2.9], 9.0, 100)
    assert 13 in idx
    # y = 2.95: .36 + 8.7025 = 9.06 > 9 -> left leaf never visited, and indeed nothing there is in range
    idx, _, _ = t.r_nearest([13.6, 2.95], 9.0, 100)
    assert list(idx) == [14]


# ---------------------------------------------------------------- (3) independent restatement
@pytest.mark.parametrize("dim,n,seed", [(2, 40, 1), (3, 40, 2), (3, 157, 3), (2, 333, 4), (3, 1000, 5)])
def test_matches_independent_python_restatement(dim, n, seed):
    rng = np.random.default_rng(seed)
    data = _pad3(rng.normal(0, 2.0, (n, dim)).astype(np.float32))
    # ties on purpose: quantise some coordinates
    data[::3, 0] = np.round(data[::3, 0])
    t = O.KdTree(data, dim)
    py = _PyTree(data[:, :dim], dim)
    d = t.dump()
    assert [i - 1 for i in d["ind"]] == py.ind                      # identical permutation
    r2 = O.search_r2()
    for _ in range(30):
        q = rng.normal(0, 2.0, dim).astype(np.float32)
        for nalloc in (3, 5, 17, n):
            idx, dis, _ = t.r_nearest(q, r2, nalloc)
            pidx, pdis = py.r_nearest(q, r2, nalloc)
            assert np.array_equal(idx, pidx)                        # same hits, same ORDER
            assert np.array_equal(dis.view(np.int32), pdis.view(np.int32))


def test_tree_shape_is_balanced_by_index():
    rng = np.random.default_rng(9)
    n = 1000
    t = O.KdTree(_pad3(rng.normal(size=(n, 3)).astype(np.float32)), 3)
    d = t.dump()
    leaves = d["left"] < 0
    assert ((d["u"] - d["l"])[leaves] <= 12).all()
    assert sorted(d["ind"]) == list(range(1, n + 1))
    internal = ~leaves
    m = (d["l"] + d["u"]) // 2
    for i in np.nonzero(internal)[0]:
        assert d["u"][d["left"][i]] == m[i] and d["l"][d["right"][i]] == m[i] + 1
