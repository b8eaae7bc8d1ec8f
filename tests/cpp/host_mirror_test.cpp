// host_mirror_test.cpp -- the reference's facade flow through the C++ host mirror (include/fspann_host.hpp) on a real GPU, written
// the way the reference's own ITs read (api/src/test/java/com/fspann/api/ForwardSecureANNSystem*Test.java, it/.../ForwardSecurityGameTest):
// construct the system, batchInsert, finalizeForSearch, createToken, search; Rotate + partial Migrate keep the results; error behaviour.
// Parity: every search result is compared bit for bit with the oracle (oracle/libfspann_oracle.so -- test infrastructure) run on
// the same index arrays and encrypted store.  Prints "HOST MIRROR OK" and exits 0 on success.
#include <cstdio>
#include <cstdlib>
#include <random>

#include "fspann_host.hpp"

extern "C" {
typedef struct {
    int32_t dim, T, D, m, lambda, W;
    int64_t N, P;
    const int64_t *min_key, *max_key;
    const uint64_t *rep;
    const int32_t *ids;
    const uint8_t *deleted;
    int64_t n_deleted_flags;
} orc_index_t;
typedef struct {
    int64_t N; int32_t dim;
    const uint8_t *iv, *ct;
    const int32_t *key_version;
    const uint8_t *deleted, *present;
    int32_t n_keys;
    const int32_t *key_versions;
    const uint8_t *keys;
} orc_store_t;
void orc_search_batch(const orc_index_t *ix, const orc_store_t *st, const double *queries, int64_t q0, int64_t q1, const double *alpha, const double *r,
                      const double *omega, int k, int probes, int64_t hard_cap, int B, int ham_threshold, int32_t *top_ids, double *top_dist, int32_t *n_ret,
                      int64_t *counters);
int orc_search(const orc_index_t *ix, const orc_store_t *st, const double *q_plain, const uint64_t *codes, int k, int probes, int64_t hard_cap, int B,
               int ham_threshold, int32_t *top_ids, double *top_dist, int32_t *cand_ids, int32_t *cand_scores, uint8_t *verdict, int64_t *counters,
               uint8_t *touched);
}

using namespace fspann;

#define CHECK(cond)                                                                 \
    do {                                                                            \
        if (!(cond)) { std::fprintf(stderr, "CHECK failed at line %d: %s\n", __LINE__, #cond); std::exit(1); } \
    } while (0)

template <class Ex, class F>
static bool throws(F f) {
    try { f(); } catch (const Ex &) { return true; } catch (...) { return false; }
    return false;
}

int main() {
    const int N = 3000, dim = 32, Q = 24, k = 10;
    SystemConfig cfg;
    cfg.paper = PaperConfig{12, 2, 4, 3, 13};          // m, lambda, divisions, tables, seed
    cfg.runtime.refinementLimit = 64;
    const int TD = cfg.paper.tables * cfg.paper.divisions, m = cfg.paper.m;
    std::mt19937_64 rng(7);
    std::normal_distribution<double> gauss(0.0, 1.0);
    // SIFT-like base: 32 cluster centres + noise, rounded to integers 0..255 (float-representable like every loader input)
    std::vector<double> centres(32 * dim), base((size_t)N * dim), queries((size_t)Q * dim);
    for (double &c : centres) c = 128.0 * (double)(rng() >> 11) * 0x1.0p-53;
    auto draw = [&](double *v) {
        const int c = (int)(rng() % 32);
        for (int i = 0; i < dim; i++) v[i] = std::rint(std::min(255.0, std::max(0.0, centres[(size_t)c * dim + i] + 20.0 * gauss(rng))));
    };
    for (int i = 0; i < N; i++) draw(&base[(size_t)i * dim]);
    for (int i = 0; i < Q; i++) draw(&queries[(size_t)i * dim]);
    // GFunctions (Coding:184-241 shape): unit-norm Gaussian rows, omega = projection range over the first 1000 vectors / 2.5, r in [0, omega)
    GFunctions g;
    g.dim = dim; g.alpha.resize((size_t)TD * m * dim); g.r.resize((size_t)TD * m); g.omega.resize((size_t)TD * m);
    for (int j = 0; j < TD * m; j++) {
        double nrm = 0;
        for (int i = 0; i < dim; i++) { g.alpha[(size_t)j * dim + i] = gauss(rng); nrm += g.alpha[(size_t)j * dim + i] * g.alpha[(size_t)j * dim + i]; }
        for (int i = 0; i < dim; i++) g.alpha[(size_t)j * dim + i] /= std::sqrt(nrm);
        double lo = 1e300, hi = -1e300;
        for (int s = 0; s < 1000; s++) {
            double y = 0;
            for (int i = 0; i < dim; i++) y += base[(size_t)s * dim + i] * g.alpha[(size_t)j * dim + i];
            lo = std::min(lo, y); hi = std::max(hi, y);
        }
        g.omega[(size_t)j] = std::max(1e-6, hi - lo) / 2.5;
        g.r[(size_t)j] = g.omega[(size_t)j] * (double)(rng() >> 11) * 0x1.0p-53;
    }
    std::vector<uint8_t> master(32);
    for (int i = 0; i < 32; i++) master[(size_t)i] = (uint8_t)(i * 7 + 1);
    uint64_t iv_ctr = 1;
    auto iv_source = [&]() {                            // deterministic stand-in for SecureRandom so the run is reproducible
        std::vector<uint8_t> iv(12);
        uint64_t x = iv_ctr++ * 0x9E3779B97F4A7C15ull;
        for (int i = 0; i < 12; i++) { x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; iv[(size_t)i] = (uint8_t)(x >> 40); }
        return iv;
    };

    ForwardSecureANNSystem sys(cfg, dim, master, g, iv_source);
    // ---- error behaviour before the index is frozen (FSA:1675-1678, PIS:594, QTF:65)
    const std::vector<double> q0(queries.begin(), queries.begin() + dim);
    CHECK(throws<IllegalStateException>([&] { sys.createToken(q0, k, dim); }));
    CHECK(throws<IllegalStateException>([&] { sys.queryService().searchBatch({QueryToken{}}); }));
    sys.batchInsert(base);
    sys.finalizeForSearch();
    CHECK(sys.index().isFrozen());
    CHECK(throws<IllegalStateException>([&] { sys.index().insert(N, q0); }));                       // "Index already finalized"
    CHECK(throws<IllegalArgumentException>([&] { sys.createToken(std::vector<double>(dim - 1, 0.0), k, dim - 1); }));
    CHECK(throws<IllegalArgumentException>([&] { sys.tokenFactory().create(q0, 0); }));
    CHECK(sys.queryService().search(nullptr).empty());                                              // QSI:102

    // ---- tokens and search
    std::vector<QueryToken> tokens;
    for (int i = 0; i < Q; i++) tokens.push_back(sys.createToken(std::vector<double>(queries.begin() + (size_t)i * dim, queries.begin() + (size_t)(i + 1) * dim), k, dim));
    CHECK(tokens[0].version == 1 && tokens[0].encryptionContext == "dim_32_v1" && tokens[0].encryptedQuery.size() == (size_t)8 * dim + 16);
    CHECK(tokens[0].iv != tokens[1].iv);                                                            // fresh IV per token
    auto res1 = sys.queryService().searchBatch(tokens);
    CHECK(sys.queryService().getLastReturned() == k && sys.queryService().getLastCandDecrypted() == 64);
    const auto cands = sys.index().lookupCandidatesWithScores(tokens[0]);
    CHECK((int)cands.size() == 64);
    for (size_t i = 1; i < cands.size(); i++) CHECK(cands[i - 1].second <= cands[i].second);        // sorted by Hamming score
    const auto touched = sys.queryService().drainTouched();
    CHECK(!touched.empty());

    // ---- parity with the oracle on the same index arrays + store (bit-exact ids and FP64 distances)
    const int64_t P = (N + 63) / 64;
    const int W = 1;
    std::vector<int64_t> mn((size_t)TD * P), mx((size_t)TD * P);
    std::vector<uint64_t> rep((size_t)TD * P * W);
    std::vector<int32_t> pids((size_t)TD * N), staged;
    for (int i = 999; i < N; i++) staged.push_back(i);
    for (int i = 0; i < 999; i++) staged.push_back(i);
    {
        GpuContext probe;
        probe.check(fspann_gfunctions_upload(probe.get(), dim, cfg.paper.tables, cfg.paper.divisions, m, cfg.paper.lambda, g.alpha.data(), g.r.data(), g.omega.data()));
        probe.check(fspann_routing_build(probe.get(), N, base.data(), staged.data(), mn.data(), mx.data(), rep.data(), pids.data()));
    }
    auto oracle_results = [&](const std::vector<int> &versions) {
        orc_index_t ix{dim, cfg.paper.tables, cfg.paper.divisions, m, cfg.paper.lambda, W, N, P, mn.data(), mx.data(), rep.data(), pids.data(), nullptr, 0};
        std::vector<uint8_t> keys;
        std::vector<int32_t> kv(versions.begin(), versions.end());
        for (int v : versions) { const auto kk = sys.keys().derive(v); keys.insert(keys.end(), kk.begin(), kk.end()); }
        orc_store_t st{N, dim, sys.store_iv.data(), sys.store_ct.data(), sys.store_ver.data(), nullptr, nullptr, (int32_t)versions.size(), kv.data(), keys.data()};
        std::vector<int32_t> ids((size_t)Q * k), nret((size_t)Q);
        std::vector<double> dist((size_t)Q * k);
        orc_search_batch(&ix, &st, queries.data(), 0, Q, g.alpha.data(), g.r.data(), g.omega.data(), k, 5, sys.index().hardCap(), cfg.runtime.refinementLimit, 0,
                         ids.data(), dist.data(), nret.data(), nullptr);
        return std::make_tuple(ids, dist, nret);
    };
    auto same_as_oracle = [&](const std::vector<std::vector<QueryResult>> &res, const std::vector<int> &versions) {
        auto [ids, dist, nret] = oracle_results(versions);
        for (int q = 0; q < Q; q++) {
            CHECK((int)res[(size_t)q].size() == nret[(size_t)q]);
            for (int r = 0; r < nret[(size_t)q]; r++) {
                CHECK(res[(size_t)q][(size_t)r].id == std::to_string(ids[(size_t)q * k + r]));
                CHECK(std::memcmp(&res[(size_t)q][(size_t)r].distance, &dist[(size_t)q * k + r], 8) == 0);
            }
        }
    };
    same_as_oracle(res1, {1});

    // ---- derive (QTF:182-200) keeps codes / IV / ciphertext; search routes on the token's OWN codes (PIS:600), never on codes
    //      recomputed from the decrypted query: a token carrying query 1's codes and query 0's ciphertext gets query 1's candidates
    {
        const QueryToken d3 = sys.tokenFactory().derive(&tokens[2], 3);
        CHECK(d3.topK == 3 && d3.bitCodes == tokens[2].bitCodes && d3.iv == tokens[2].iv && d3.encryptedQuery == tokens[2].encryptedQuery);
        const auto r3 = sys.queryService().search(&d3);
        CHECK(r3.size() == 3 && r3[0].id == res1[2][0].id && r3[2].id == res1[2][2].id);
        CHECK(throws<IllegalArgumentException>([&] { sys.tokenFactory().derive(&tokens[2], 0); }));
        CHECK(throws<IllegalArgumentException>([&] { sys.tokenFactory().derive(nullptr, 3); }));
        QueryToken sw = tokens[0];
        sw.bitCodes = tokens[1].bitCodes;
        const auto rs = sys.queryService().search(&sw);
        orc_index_t ix{dim, cfg.paper.tables, cfg.paper.divisions, m, cfg.paper.lambda, W, N, P, mn.data(), mx.data(), rep.data(), pids.data(), nullptr, 0};
        const auto k1 = sys.keys().derive(1);
        const int32_t kv1 = 1;
        orc_store_t st{N, dim, sys.store_iv.data(), sys.store_ct.data(), sys.store_ver.data(), nullptr, nullptr, 1, &kv1, k1.data()};
        std::vector<int32_t> oid((size_t)k), cid(65), csc(65);
        std::vector<double> od((size_t)k);
        std::vector<uint8_t> ver(65);
        int64_t cnt[6];
        const int n = orc_search(&ix, &st, queries.data(), tokens[1].bitCodes.data(), k, 5, sys.index().hardCap(), cfg.runtime.refinementLimit, 0, oid.data(),
                                 od.data(), cid.data(), csc.data(), ver.data(), cnt, nullptr);
        CHECK((int)rs.size() == n);
        for (int r = 0; r < n; r++) CHECK(rs[(size_t)r].id == std::to_string(oid[(size_t)r]) && std::memcmp(&rs[(size_t)r].distance, &od[(size_t)r], 8) == 0);
        QueryToken nocodes = tokens[0];
        nocodes.bitCodes.clear();
        CHECK(throws<IllegalStateException>([&] { sys.queryService().search(&nocodes); }));       // PIS:604-606
    }

    // ---- Rotate -> v2, Migrate ids = 0 (mod 3) on the device: results must not move; Retire v1 is refused while records are bound
    CHECK(sys.rotateKeyOnly() == 2);
    std::vector<int32_t> mig;
    for (int i = 0; i < N; i += 3) mig.push_back(i);
    const std::vector<uint8_t> ct_before = sys.store_ct;
    CHECK(sys.reencryptTouched(mig, 2) == (int64_t)mig.size());
    CHECK(sys.store_ct != ct_before && sys.store_ver[0] == 2 && sys.store_ver[1] == 1);
    auto res2 = sys.queryService().searchBatch(tokens);                                             // v1 tokens stay decryptable (QSI:124-129)
    same_as_oracle(res2, {1, 2});
    for (int q = 0; q < Q; q++) for (size_t r = 0; r < res1[(size_t)q].size(); r++) CHECK(res1[(size_t)q][r].id == res2[(size_t)q][r].id && res1[(size_t)q][r].distance == res2[(size_t)q][r].distance);
    CHECK(!sys.retire(1));
    CHECK(sys.createToken(q0, k, dim).version == 2);
    // migrate the rest, retire v1, fresh tokens under v2: same results again
    std::vector<int32_t> rest;
    for (int i = 0; i < N; i++) if (sys.store_ver[(size_t)i] == 1) rest.push_back(i);
    CHECK(sys.reencryptTouched(rest, 2) == (int64_t)rest.size());
    CHECK(sys.reencryptTouched(rest, 2) == 0);                                                      // already upgraded -> skipped (KRS:248)
    CHECK(sys.retire(1));
    std::vector<QueryToken> tokens2;
    for (int i = 0; i < Q; i++) tokens2.push_back(sys.createToken(std::vector<double>(queries.begin() + (size_t)i * dim, queries.begin() + (size_t)(i + 1) * dim), k, dim));
    auto res3 = sys.queryService().searchBatch(tokens2);
    same_as_oracle(res3, {2});
    for (int q = 0; q < Q; q++) CHECK(res3[(size_t)q].size() == res1[(size_t)q].size() && res3[(size_t)q][0].id == res1[(size_t)q][0].id);
    // a token whose key version is gone falls back to the current key and fails authentication -> "Query decryption failed" (AGC:199-203)
    CHECK(throws<std::runtime_error>([&] { sys.queryService().searchBatch({tokens[0]}); }));
    // probe override is consumed by one search (QSI:343)
    sys.index().setProbeOverride(8);
    CHECK(sys.index().effectiveMaxProbes() == 8);
    sys.queryService().searchBatch({tokens2[0]});
    CHECK(sys.index().effectiveMaxProbes() == 5);
    std::printf("HOST MIRROR OK: %d queries x 3 key states bit-exact vs the oracle, %zu + %zu records migrated on the device\n", Q, mig.size(), rest.size());
    return 0;
}
