"""Independent pure-Python re-implementations of the JDK behaviours the reference's results depend on, checked against the
oracle: java.util.HashMap iteration order (with real bucket arrays and resize splits), java.util.PriorityQueue sift rules,
and a literal re-statement of lookupCandidatesWithScores (PIS:592-715) on tiny random indexes."""
import numpy as np
import pytest

from oracle import oracle as O


def jhash(s: str) -> int:
    h = 0
    for ch in s:
        h = (31 * h + ord(ch)) & 0xFFFFFFFF
    return h ^ (h >> 16)


class JavaHashMap:
    """java.util.HashMap<String, V> with linked bins only (no treeification), including resize()'s lo/hi split."""

    def __init__(self, initial_capacity):
        n = 1
        while n < initial_capacity:
            n <<= 1
        self.threshold0 = n if initial_capacity > 0 else 1
        self.table = None
        self.size = 0
        self.max_chain = 0

    def _resize(self):
        if self.table is None:
            self.table = [[] for _ in range(self.threshold0)]
        else:
            old = self.table
            new = [[] for _ in range(2 * len(old))]
            for j, b in enumerate(old):
                for e in b:                       # relative order preserved inside lo and hi lists
                    new[j + len(old) if (e[0] & len(old)) else j].append(e)
            self.table = new
        self.thr = int(0.75 * len(self.table))

    def put(self, key, value):
        if self.table is None:
            self._resize()
        h = jhash(key)
        b = self.table[h & (len(self.table) - 1)]
        for e in b:
            if e[1] == key:
                e[2] = value
                return
        b.append([h, key, value])
        self.max_chain = max(self.max_chain, len(b))
        self.size += 1
        if self.size > self.thr:
            self._resize()

    def get(self, key):
        if self.table is None:
            return None
        for e in self.table[jhash(key) & (len(self.table) - 1)]:
            if e[1] == key:
                return e[2]
        return None

    def items(self):
        if self.table is None:
            return
        for b in self.table:
            for e in b:
                yield e[1], e[2]


@pytest.mark.parametrize("n,cap", [(1, 1), (13, 16), (100, 16), (1000, 1000), (5000, 64), (3000, 4096), (20000, 20000)])
def test_hashmap_iteration_order(n, cap):
    rng = np.random.default_rng(n + cap)
    keys = rng.permutation(max(4 * n, 50))[:n].astype(np.int32)
    m = JavaHashMap(cap)
    for k in keys:
        m.put(str(int(k)), 1)
    order, max_chain = O.hashmap_order(keys, cap)
    assert [int(keys[i]) for i in order] == [int(k) for k, _ in m.items()]
    assert max_chain == m.max_chain
    assert all(O.java_hash_decimal(int(k)) == jhash(str(int(k))) for k in keys[:200])


class JavaPQ:
    """java.util.PriorityQueue with a comparator on entry[1] (siftUp / siftDown exactly as in the JDK)."""

    def __init__(self):
        self.q = []

    def add(self, x):
        k = len(self.q)
        self.q.append(x)
        while k > 0:
            parent = (k - 1) >> 1
            if x[1] >= self.q[parent][1]:
                break
            self.q[k] = self.q[parent]
            k = parent
        self.q[k] = x

    def poll(self):
        res = self.q[0]
        x = self.q.pop()
        n = len(self.q)
        if n > 0:
            k, half = 0, n >> 1
            while k < half:
                child = 2 * k + 1
                c = self.q[child]
                if child + 1 < n and c[1] > self.q[child + 1][1]:
                    child += 1
                    c = self.q[child]
                if x[1] <= c[1]:
                    break
                self.q[k] = c
                k = child
            self.q[k] = x
        return res


def py_route(ix: O.Index, codes, probes, hard_cap):
    """Literal restatement of PIS:592-715 + 726-753 in Python on top of JavaHashMap / JavaPQ."""
    g = ix.g
    TD, W, P, N = g.T * g.D, g.W, ix.P, ix.N
    best = JavaHashMap(min(hard_cap, 1 << 16))
    raw = 0
    for td in range(TD):
        if best.size >= hard_cap:
            break
        q = codes[td]
        qkey = O.compute_key(q)
        mn, mx = ix.min_key[td], ix.max_key[td]
        lo, hi, center = 0, P - 1, None
        while lo <= hi:
            mid = (lo + hi) >> 1
            if qkey < mn[mid]:
                hi = mid - 1
            elif qkey > mx[mid]:
                lo = mid + 1
            else:
                center = mid
                break
        if center is None:
            if lo <= 0:
                center = 0
            elif lo >= P:
                center = P - 1
            else:
                def dist(i):
                    return mn[i] - qkey if qkey < mn[i] else (qkey - mx[i] if qkey > mx[i] else 0)
                center = lo - 1 if dist(lo - 1) <= dist(lo) else lo

        def ham(i):
            return sum(bin(int(a) ^ int(b)).count("1") for a, b in zip(q, ix.rep[td, i]))
        pq = JavaPQ()
        visited = set([center])
        pq.add((center, ham(center)))
        used = 0
        while pq.q and used < probes and best.size < hard_cap:
            idx, _ = pq.poll()
            used += 1
            pd = ham(idx)
            for j in range(idx * 64, min(idx * 64 + 64, N)):
                id_ = str(int(ix.ids[td, j]))
                prev = best.get(id_)
                if prev is None or pd < prev:
                    best.put(id_, pd)
                    raw += 1
            for nb in (idx - 1, idx + 1):
                if 0 <= nb < P and nb not in visited:
                    visited.add(nb)
                    pq.add((nb, ham(nb)))
    entries = list(best.items())
    entries.sort(key=lambda kv: kv[1])          # list.sort is stable, like java.util.List.sort
    return [int(k) for k, _ in entries], [v for _, v in entries], raw, best.max_chain


@pytest.mark.parametrize("seed,probes,hard_cap", [(1, 5, 20000), (2, 3, 300), (3, 10, 20000), (4, 1, 50), (5, 5, 1000)])
def test_route_matches_literal_python_restatement(seed, probes, hard_cap):
    rng = np.random.default_rng(seed)
    N, dim, T, D, m, lam = 1300, 12, 2, 3, 10, 2
    base = rng.normal(size=(N, dim)).astype(np.float32).astype(np.float64)
    g = O.registry_init(base[:1000], m, lam, 13, T, D)
    codes = O.tokengen_batch(base, g)
    ix = O.index_build(codes, g, O.staged_order(N))
    qs = rng.normal(size=(6, dim)).astype(np.float32).astype(np.float64)
    qc = O.tokengen_batch(qs, g)
    for q in range(qs.shape[0]):
        ids, sc, raw, mc = O.route(ix, qc[q], probes, hard_cap)
        pids, psc, praw, pmc = py_route(ix, qc[q], probes, hard_cap)
        assert raw == praw and list(sc) == psc
        if pmc < 9:
            assert list(ids) == pids
        else:
            assert sorted(ids.tolist()) == sorted(pids)


def test_partition_build_matches_python_restatement():
    """GP.build (GP:37-76) over HashMap<String,BitSet>(N) iteration order (PIS:412-420)."""
    rng = np.random.default_rng(11)
    N, dim, T, D, m, lam = 1500, 8, 1, 2, 9, 3
    base = rng.normal(size=(N, dim)).astype(np.float32).astype(np.float64)
    g = O.registry_init(base[:1000], m, lam, 7, T, D)
    codes = O.tokengen_batch(base, g)
    staged = O.staged_order(N)
    assert staged[0] == 999 and staged[-1] == 998 and len(set(staged.tolist())) == N
    ix = O.index_build(codes, g, staged)
    for td in range(T * D):
        hm = JavaHashMap(N)
        for i in staged:
            hm.put(str(int(i)), codes[i, td])
        ordered = [(int(k), O.compute_key(c)) for k, c in hm.items()]
        ordered.sort(key=lambda kv: kv[1])
        assert [k for k, _ in ordered] == ix.ids[td].tolist()
        for p in range(ix.P):
            blk = ordered[p * 64:(p + 1) * 64]
            assert ix.min_key[td, p] == blk[0][1] and ix.max_key[td, p] == blk[-1][1]
            mid = (len(blk) - 1) >> 1
            assert np.array_equal(ix.rep[td, p], codes[blk[mid][0], td])


def test_coding_bit_layout():
    """Coding.C (Coding:285-301): bit (lambda-1-i)*m + j = bit i of H[j]; CodingQuickCheck: bit 0 of C(v) = MSB of H[0]."""
    rng = np.random.default_rng(3)
    base = rng.normal(size=(1000, 16)).astype(np.float32).astype(np.float64)
    for m, lam in [(24, 2), (10, 3), (22, 2), (24, 3)]:
        g = O.registry_init(base, m, lam, 13, 2, 2)
        codes = O.tokengen_batch(base[:50], g)
        for v in range(50):
            for td in range(4):
                Hv = O.H(base[v], g, td)
                bits = np.zeros(m * lam, dtype=np.uint8)
                for i in range(lam):
                    for j in range(m):
                        bits[(lam - 1 - i) * m + j] = (int(Hv[j]) >> i) & 1
                got = np.unpackbits(codes[v, td].view(np.uint8), bitorder="little")[: m * lam]
                assert np.array_equal(got, bits)
                assert got[0] == (int(Hv[0]) >> (lam - 1)) & 1
                # H itself: floor((v.alpha + r)/omega), sequential FP64
                y = 0.0
                for i in range(16):
                    y += base[v, i] * g.alpha[td, 0, i]
                assert Hv[0] == int(np.floor((y + g.r[td, 0]) / g.omega[td, 0]))
