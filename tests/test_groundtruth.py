"""exact ground truth (GroundtruthPrecompute.run) and recall@K (FSA:785-794).
CPU part: the oracle against an independent numpy restatement.  GPU part: the device kernels against the oracle, bit-exact
ids AND squared distances, including heavy ties (integer SIFT-like data), K = 1 / K = N, odd dimensions."""
import numpy as np
import pytest

from oracle import oracle as O


def numpy_gt(base, q, K):
    ids = np.empty((q.shape[0], K), dtype=np.int32)
    d2 = np.empty((q.shape[0], K), dtype=np.float64)
    for i in range(q.shape[0]):
        diff = (q[i][None, :] - base).astype(np.float32).astype(np.float64)      # float subtraction, widened (GTP:149-151)
        s = np.zeros(base.shape[0])
        for j in range(base.shape[1]):                                            # sequential FP64 sum
            s = s + diff[:, j] * diff[:, j]
        o = np.lexsort((np.arange(base.shape[0]), s))[:K]
        ids[i], d2[i] = o, s[o]
    return ids, d2


def datasets():
    rng = np.random.default_rng(5)
    yield "sift-int-ties", rng.integers(0, 4, size=(3000, 8)).astype(np.float32), rng.integers(0, 4, size=(9, 8)).astype(np.float32), 50
    yield "gauss-33", rng.normal(0, 1, size=(2500, 33)).astype(np.float32), rng.normal(0, 1, size=(70, 33)).astype(np.float32), 10
    yield "sift128", rng.integers(0, 256, size=(5000, 128)).astype(np.float32), rng.integers(0, 256, size=(65, 128)).astype(np.float32), 100
    yield "k-equals-n", rng.normal(0, 1, size=(300, 20)).astype(np.float32), rng.normal(0, 1, size=(5, 20)).astype(np.float32), 300
    yield "k1", rng.normal(0, 1, size=(1000, 4)).astype(np.float32), rng.normal(0, 1, size=(3, 4)).astype(np.float32), 1
    same = np.tile(rng.normal(0, 1, size=(1, 12)).astype(np.float32), (6000, 1))   # every base vector identical: 6000 exact ties
    yield "all-tied", same, rng.normal(0, 1, size=(4, 12)).astype(np.float32), 17


@pytest.mark.parametrize("name,base,q,K", list(datasets()), ids=[d[0] for d in datasets()])
def test_oracle_groundtruth_matches_numpy_restatement(name, base, q, K):
    ids, d2 = O.groundtruth(base, q, K)
    rid, rd2 = numpy_gt(base, q, K)
    assert np.array_equal(ids, rid)
    assert np.array_equal(d2.view(np.uint64), rd2.view(np.uint64))


def test_oracle_recall_at_k():
    gt = np.arange(10, dtype=np.int32)
    assert O.recall_at_k(gt, gt[::-1].copy(), 10, 10) == 1.0
    assert O.recall_at_k(gt, np.array([3, 99, 4, 98, 97, 96, 95, 94, 93, 0], dtype=np.int32), 10, 10) == 0.3
    assert O.recall_at_k(gt, np.array([3, 4, 5, 6, 7, 8, 9, 0, 1, 2], dtype=np.int32), 4, 10) == 0.4     # only n_ret results count
    assert O.recall_at_k(gt, np.array([12, 11, 10, 0], dtype=np.int32), 4, 3) == 0.0                     # only the first K count


@pytest.mark.gpu
@pytest.mark.parametrize("name,base,q,K", list(datasets()), ids=[d[0] for d in datasets()])
def test_gpu_groundtruth_bit_exact(name, base, q, K):
    from fspann_query_system_b200.gpu import GpuContext
    ctx = GpuContext(0)
    try:
        ids, d2 = ctx.groundtruth(base, q, K, want_d2=True)
        rid, rd2 = O.groundtruth(base, q, K)
        assert np.array_equal(ids, rid)
        assert np.array_equal(d2.view(np.uint64), rd2.view(np.uint64))
    finally:
        ctx.close()


@pytest.mark.gpu
def test_gpu_recall_and_end_to_end_recall_of_search(world_factory):
    """recall@10 of the GPU search against GPU ground truth equals the same computed with the oracle's search and ground truth."""
    from fspann_query_system_b200 import _native as N
    from fspann_query_system_b200.gpu import GpuContext
    w = world_factory(N=5000, dim=128, Q=48, T=4, D=8, m=24, lam=2)
    ctx = w.gpu_context()
    try:
        K = 10
        gt = ctx.groundtruth(w.base.astype(np.float32), w.queries.astype(np.float32), K)
        gt_ref, _ = O.groundtruth(w.base.astype(np.float32), w.queries.astype(np.float32), K)
        assert np.array_equal(gt, gt_ref)
        got = ctx.search_batch(w.queries, K, 5, 20000, 256)
        rec = ctx.recall_batch(gt, got["top_ids"], K, got["n_ret"])
        ref = np.array([O.recall_at_k(gt_ref[q], got["top_ids"][q], int(got["n_ret"][q]), K) for q in range(w.queries.shape[0])])
        assert np.array_equal(rec, ref)
        assert 0.0 < rec.mean() <= 1.0
        # n_ret = None means K results per row; a truncated row only counts what was returned
        assert np.array_equal(ctx.recall_batch(gt, gt, K), np.ones(gt.shape[0]))
        half = np.full(gt.shape[0], 5, dtype=np.int32)
        assert np.array_equal(ctx.recall_batch(gt, gt, K, half), np.full(gt.shape[0], 0.5))
        with pytest.raises(N.IllegalArgumentError):
            ctx.recall_batch(gt[:, :5], gt, K)                        # groundtruth shorter than K (FSA:779-782)
        with pytest.raises(N.IllegalArgumentError):
            ctx.groundtruth(w.base.astype(np.float32), w.queries.astype(np.float32)[:, :64], K)
    finally:
        ctx.close()
