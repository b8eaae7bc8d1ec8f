"""Golden vectors produced by the REAL reference (integration/java/com/fspann/gpu/GoldenDump.java run against the stock Java code).

No JVM exists in this image, so tests/golden/reference/ is absent here and these tests SKIP LOUDLY: parity stays "unpinned" until a
maintainer with a JDK runs GoldenDump once and commits its output directory.  From then on:
  * CPU (-m "not gpu"): the oracle must reproduce every Java intermediate bit for bit -- routing codes, ordered candidate lists + raw
    counts, AES-GCM records (decrypting to the dumped base vectors), top-k ids, FP64 distances and the getLast* counters;
  * GPU (-m gpu): the CUDA path must reproduce the same through the C ABI.
`test_consumer_self_check` runs the very same checker on a directory written by the ORACLE in GoldenDump's format, so the consumer
code is exercised on every run; it pins nothing (oracle vs itself) and says so."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "golden", "reference")
SKIP_MSG = ("tests/golden/reference/manifest.json is absent: parity with the real Java reference is UNPINNED.  Run "
            "integration/java/com/fspann/gpu/GoldenDump.java against the stock reference (needs a JDK) and commit its output.")


def load_dump(path):
    man = json.load(open(os.path.join(path, "manifest.json")))
    arrays = {}
    for name, meta in man["arrays"].items():
        a = np.fromfile(os.path.join(path, name + ".bin"), dtype=np.dtype(meta["dtype"]))
        arrays[name] = a.reshape(meta["shape"])
    return man["params"], arrays


def write_dump(path, params, arrays):
    """GoldenDump's on-disk format (little-endian raw arrays + manifest.json), used by the self-check."""
    os.makedirs(path, exist_ok=True)
    man = {"params": params, "arrays": {}}
    for name, a in arrays.items():
        a = np.ascontiguousarray(a)
        a2 = a.reshape(a.shape[0], -1) if a.ndim != 2 else a
        a2.tofile(os.path.join(path, name + ".bin"))
        man["arrays"][name] = {"dtype": a2.dtype.str, "shape": list(a2.shape)}
    json.dump(man, open(os.path.join(path, "manifest.json"), "w"))


def build_oracle_objects(p, a):
    T, D, m, lam, dim, N = p["tables"], p["divisions"], p["m"], p["lambda"], p["dim"], p["N"]
    g = O.GFunctions(dim, T, D, m, lam, a["alpha"].reshape(T * D, m, dim), a["r"].reshape(T * D, m), a["omega"].reshape(T * D, m))
    base = a["base"]
    codes_base = O.tokengen_batch(base, g)
    ix = O.index_build(codes_base, g, O.staged_order(N))
    keys = {int(v): bytes(a["keys"][i]) for i, v in enumerate(a["key_versions"].ravel())}
    store = O.Store(dim, a["store_iv"], a["store_ct"], a["store_key_version"].ravel().astype(np.int32), keys)
    return g, ix, store


def check_against_dump(p, a, engine=None):
    """Every intermediate of the dump vs the oracle (engine=None) or vs a GpuContext holding the same state (engine=ctx)."""
    g, ix, store = build_oracle_objects(p, a)
    Q, k, B = p["Q"], p["k"], p["refinementLimit"]
    probes, hard_cap = p["probes"], max(p["maxGlobalCandidates"], B)
    W = (p["m"] * p["lambda"] + 63) // 64
    queries = a["queries"]
    want_codes = a["codes"].reshape(Q, p["tables"] * p["divisions"], W)
    got_codes = O.tokengen_batch(queries, g) if engine is None else engine.tokengen_batch(queries)
    assert np.array_equal(got_codes, want_codes), "routing codes differ from the reference's QueryToken.getBitCodes()"
    # the store decrypts to the dumped base vectors under the dumped keys with the reference's AAD (AGC:126-166, EP:80-83)
    sample = np.arange(0, p["N"], max(1, p["N"] // 97), dtype=np.int32)
    ref = O.refine(store, queries[0], sample, len(sample), want_plaintext=True)
    assert (ref["verdict"] == 0).all() and np.array_equal(ref["plaintext"], a["base"][sample])
    off = np.concatenate([[0], np.cumsum(a["cand_count"].ravel())])
    if engine is not None:
        out = engine.search_tokens(want_codes, queries, k, probes, hard_cap, B)
    for q in range(Q):
        want_ids, want_sc = a["cand_ids"].ravel()[off[q]:off[q + 1]], a["cand_scores"].ravel()[off[q]:off[q + 1]]
        ids, sc, raw, _ = O.route(ix, want_codes[q], probes, hard_cap)
        assert np.array_equal(ids, want_ids) and np.array_equal(sc, want_sc), f"query {q}: ordered candidate list differs (PIS:592-715)"
        assert raw == int(a["cand_raw_count"].ravel()[q])
        n = int(a["n_ret"].ravel()[q])
        if engine is None:
            r = O.search(ix, store, queries[q], want_codes[q], k, probes, hard_cap, B)
            got_ids, got_d, cnt = r["top_ids"], r["top_dist"], (r["cand_total"], r["cand_kept"], r["cand_decrypted"], r["returned"])
        else:
            got_ids, got_d = out["top_ids"][q, :out["n_ret"][q]], out["top_dist"][q, :out["n_ret"][q]]
            cnt = tuple(int(x) for x in out["counters"][q, :4])
        assert len(got_ids) == n and np.array_equal(got_ids, a["top_ids"][q, :n]), f"query {q}: top-k ids differ (QSI:298-316)"
        assert np.array_equal(np.asarray(got_d).view(np.uint64), a["top_dist"][q, :n].view(np.uint64)), f"query {q}: FP64 distances differ"
        assert cnt == tuple(int(x) for x in a["counters"][q]), f"query {q}: getLast* counters differ"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DIR, "manifest.json")), reason=SKIP_MSG)
def test_oracle_matches_reference_goldens():
    p, a = load_dump(REF_DIR)
    check_against_dump(p, a)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DIR, "manifest.json")), reason=SKIP_MSG)
def test_cuda_path_matches_reference_goldens():
    from fspann_query_system_b200.gpu import GpuContext
    p, a = load_dump(REF_DIR)
    g, ix, store = build_oracle_objects(p, a)
    ctx = GpuContext(0)
    try:
        ctx.routing_upload(g.dim, g.T, g.D, g.m, g.lam, g.alpha, g.r, g.omega, ix.min_key, ix.max_key, ix.rep, ix.ids)
        for v, key in store.keys.items():
            ctx.keys_set(v, key)
        ctx.store_upload(g.dim, store.iv, store.ct, store.key_version)
        check_against_dump(p, a, engine=ctx)
    finally:
        ctx.close()


def test_consumer_self_check(tmp_path, world_factory):
    """The checker above on a dump written by the ORACLE in GoldenDump's format: exercises loader + comparisons (pins nothing)."""
    w = world_factory(N=3000, dim=32, Q=40, T=3, D=4, m=12, lam=2)
    g, Q, k, B = w.g, w.queries.shape[0], 10, 64
    codes = O.tokengen_batch(w.queries, g)
    cand_ids, cand_sc, cnt, raw = [], [], [], []
    top_ids = np.full((Q, k), -1, np.int32); top_d = np.zeros((Q, k)); n_ret = np.zeros(Q, np.int32); counters = np.zeros((Q, 4), np.int32)
    for q in range(Q):
        ids, sc, rw, _ = O.route(w.ix, codes[q], 5, 20000)
        cand_ids.append(ids); cand_sc.append(sc); cnt.append(len(ids)); raw.append(rw)
        r = O.search(w.ix, w.store, w.queries[q], codes[q], k, 5, 20000, B)
        n = len(r["top_ids"]); n_ret[q] = n; top_ids[q, :n] = r["top_ids"]; top_d[q, :n] = r["top_dist"]
        counters[q] = (r["cand_total"], r["cand_kept"], r["cand_decrypted"], r["returned"])
    params = dict(N=w.cfg.N, dim=g.dim, Q=Q, k=k, m=g.m, tables=g.T, divisions=g.D, seed=13, refinementLimit=B, maxGlobalCandidates=20000, probes=5)
    params["lambda"] = g.lam
    arrays = dict(base=w.base, queries=w.queries, alpha=g.alpha.reshape(-1, g.dim), r=g.r, omega=g.omega, store_iv=w.store.iv, store_ct=w.store.ct,
                  store_key_version=w.store.key_version.astype("<i4"), key_versions=np.array(sorted(w.store.keys), dtype="<i4"),
                  keys=np.stack([np.frombuffer(w.store.keys[v], dtype=np.uint8) for v in sorted(w.store.keys)]), codes=codes.reshape(Q, -1),
                  cand_count=np.array(cnt, dtype="<i4"), cand_raw_count=np.array(raw, dtype="<i4"), cand_ids=np.concatenate(cand_ids).astype("<i4"),
                  cand_scores=np.concatenate(cand_sc).astype("<i4"), top_ids=top_ids, top_dist=top_d, n_ret=n_ret, counters=counters)
    write_dump(str(tmp_path), params, arrays)
    p, a = load_dump(str(tmp_path))
    check_against_dump(p, a)
    a["top_ids"][3, 0] += 1                                     # and the checker does notice a difference
    with pytest.raises(AssertionError):
        check_against_dump(p, a)
