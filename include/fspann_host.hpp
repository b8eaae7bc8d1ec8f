// fspann_host.hpp -- C++17 host-side mirror of the reference's operator interface for the query hot path, above the C ABI of
// libfspann_gpu.so (include/fspann_gpu.h).  Header only; needs OpenSSL's libcrypto for the host-side pieces the reference keeps on
// the host (query-token AES-GCM, HMAC-SHA256 key derivation).
//
// The reference is Java and no JDK exists in this image, so this is the compiled-language stand-in for the classes a Java host
// would keep (INTEGRATION.md shows the JNI / Panama binding of the same ABI).  Same names, argument meaning and error behaviour:
//   SystemConfig                config/src/main/java/com/fspann/config/SystemConfig.java:44-85,237-337
//   QueryToken / QueryResult    common/src/main/java/com/fspann/common/QueryToken.java:23-71, QueryResult.java:6-23
//   KeyManager                  keymanagement/src/main/java/com/fspann/key/KeyManager.java:125-153,221-237,274-317
//   QueryTokenFactory.create    query/src/main/java/com/fspann/query/core/QueryTokenFactory.java:63-167
//   PartitionedIndexService     index/src/main/java/com/fspann/index/paper/PartitionedIndexService.java (insert 266, finalizeForSearch 789,
//                               lookupCandidatesWithScores 592, set/clearProbeOverride 868-874, effective probes 880-888)
//   QueryServiceImpl.search     query/src/main/java/com/fspann/query/service/QueryServiceImpl.java:100-352
//   ForwardSecureANNSystem      api/src/main/java/com/fspann/api/ForwardSecureANNSystem.java (batchInsert 479, finalizeForSearch 977,
//                               createToken 1673), KeyRotationServiceImpl.reencryptTouched / rotateKeyOnly (KRS:215-298)
// IllegalArgumentException / IllegalStateException are thrown where the Java code throws them; the ABI's FSPANN_E_ARG / FSPANN_E_STATE
// map onto the same two classes.  TokenGen, Route, Refine, bulk encryption, the index build and Migrate run on the GPU.
#pragma once
#include <openssl/evp.h>
#include <openssl/hmac.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "fspann_gpu.h"

namespace fspann {

struct IllegalArgumentException : std::invalid_argument { using std::invalid_argument::invalid_argument; };
struct IllegalStateException : std::logic_error { using std::logic_error::logic_error; };

struct PaperConfig { int m = 24, lambda = 2, divisions = 8, tables = 8; long long seed = 13; };                  // CFG:237-263
struct RuntimeConfig { int refinementLimit = 1024; long long maxGlobalCandidates = 20000; int probeOverride = -1; int hammingPrefilterThreshold = 0; };
struct SystemConfig { PaperConfig paper; RuntimeConfig runtime; };

struct QueryResult { std::string id; double distance; };                                                        // QueryResult.java:6-23
struct QueryToken {                                                                                             // QT:28-44
    std::vector<uint64_t> bitCodes;   // [tables][divisions][W] BitSet.toLongArray words
    std::vector<uint8_t> iv, encryptedQuery;
    int topK = 0, numTables = 0, dimension = 0, version = 0, lambda = 0;
    std::string encryptionContext;
};

// ---------------------------------------------------------------- host-side crypto the reference keeps on the host
class KeyManager {                    // K_v = HMAC-SHA256(K_M, be32(v))[:32] (KM:221-237)
   public:
    explicit KeyManager(std::vector<uint8_t> master) : master_(std::move(master)) {
        if (master_.size() != 32) throw IllegalArgumentException("master key must be 32 bytes");
        live_.insert(1);
    }
    std::vector<uint8_t> derive(int version) const {
        uint8_t msg[4] = {(uint8_t)(version >> 24), (uint8_t)(version >> 16), (uint8_t)(version >> 8), (uint8_t)version}, out[32];
        unsigned len = 32;
        HMAC(EVP_sha256(), master_.data(), (int)master_.size(), msg, 4, out, &len);
        return std::vector<uint8_t>(out, out + 32);
    }
    std::vector<uint8_t> getVersion(int version) const {                                                        // KRS:82-88
        if (!live_.count(version)) throw IllegalArgumentException("Unknown key version: " + std::to_string(version));
        return derive(version);
    }
    int getCurrentVersion() const { return current_; }
    int rotateKey() { live_.insert(++current_); return current_; }                                              // KM:133-153
    void retire(int version) { live_.erase(version); }                                                          // KM:274-317
    bool isLive(int version) const { return live_.count(version) != 0; }

   private:
    std::vector<uint8_t> master_;
    int current_ = 1;
    std::set<int> live_;
};

namespace detail {
inline std::vector<uint8_t> be_doubles(const double *v, int n) {                                                // AGC:240-259
    std::vector<uint8_t> out((size_t)n * 8);
    for (int i = 0; i < n; i++) {
        uint64_t b;
        std::memcpy(&b, &v[i], 8);
        for (int k = 0; k < 8; k++) out[(size_t)i * 8 + k] = (uint8_t)(b >> (56 - 8 * k));
    }
    return out;
}
// AES-256-GCM without AAD (AGC:169-204): encryptQuery / decryptQuery
inline std::vector<uint8_t> gcm_encrypt(const std::vector<uint8_t> &key, const std::vector<uint8_t> &iv, const std::vector<uint8_t> &pt) {
    std::vector<uint8_t> out(pt.size() + 16);
    EVP_CIPHER_CTX *c = EVP_CIPHER_CTX_new();
    int len = 0, ok = EVP_EncryptInit_ex(c, EVP_aes_256_gcm(), nullptr, nullptr, nullptr) && EVP_CIPHER_CTX_ctrl(c, EVP_CTRL_GCM_SET_IVLEN, (int)iv.size(), nullptr) &&
                      EVP_EncryptInit_ex(c, nullptr, nullptr, key.data(), iv.data()) && EVP_EncryptUpdate(c, out.data(), &len, pt.data(), (int)pt.size());
    int fin = 0;
    ok = ok && EVP_EncryptFinal_ex(c, out.data() + len, &fin) && EVP_CIPHER_CTX_ctrl(c, EVP_CTRL_GCM_GET_TAG, 16, out.data() + pt.size());
    EVP_CIPHER_CTX_free(c);
    if (!ok) throw std::runtime_error("Query encryption failed");
    return out;
}
inline std::vector<double> gcm_decrypt_doubles(const std::vector<uint8_t> &key, const std::vector<uint8_t> &iv, const std::vector<uint8_t> &ct) {
    if (ct.size() < 16 || (ct.size() - 16) % 8) throw std::runtime_error("Query decryption failed");
    const size_t n = ct.size() - 16;
    std::vector<uint8_t> pt(n);
    EVP_CIPHER_CTX *c = EVP_CIPHER_CTX_new();
    int len = 0, fin = 0;
    int ok = EVP_DecryptInit_ex(c, EVP_aes_256_gcm(), nullptr, nullptr, nullptr) && EVP_CIPHER_CTX_ctrl(c, EVP_CTRL_GCM_SET_IVLEN, (int)iv.size(), nullptr) &&
             EVP_DecryptInit_ex(c, nullptr, nullptr, key.data(), iv.data()) && EVP_DecryptUpdate(c, pt.data(), &len, ct.data(), (int)n) &&
             EVP_CIPHER_CTX_ctrl(c, EVP_CTRL_GCM_SET_TAG, 16, const_cast<uint8_t *>(ct.data() + n)) && EVP_DecryptFinal_ex(c, pt.data() + len, &fin) > 0;
    EVP_CIPHER_CTX_free(c);
    if (!ok) throw std::runtime_error("Query decryption failed");                                               // AGC:199-203
    std::vector<double> v(n / 8);
    for (size_t i = 0; i < v.size(); i++) {
        uint64_t b = 0;
        for (int k = 0; k < 8; k++) b = (b << 8) | pt[i * 8 + k];
        std::memcpy(&v[i], &b, 8);
    }
    return v;
}
}  // namespace detail

// ---------------------------------------------------------------- RAII over the C ABI
class GpuContext {
   public:
    explicit GpuContext(int device = 0) {
        if (fspann_ctx_create(device, &ctx_) != FSPANN_OK) throw std::runtime_error("no CUDA device: the FSPANN hot path has no CPU fallback");
    }
    ~GpuContext() { if (ctx_) fspann_ctx_destroy(ctx_); }
    GpuContext(const GpuContext &) = delete;
    GpuContext &operator=(const GpuContext &) = delete;
    fspann_ctx *get() const { return ctx_; }
    void check(int rc) const {
        if (rc == FSPANN_OK) return;
        const std::string msg = fspann_last_error(ctx_);
        if (rc == FSPANN_E_ARG) throw IllegalArgumentException(msg);
        if (rc == FSPANN_E_STATE) throw IllegalStateException(msg);
        throw std::runtime_error("CUDA failure " + std::to_string(rc) + ": " + msg);
    }

   private:
    fspann_ctx *ctx_ = nullptr;
};

// GFunctionRegistry contents cross the boundary as data (GFR:63-147; Java's Math.log/cos are not reproducible elsewhere).
struct GFunctions { int dim = 0; std::vector<double> alpha, r, omega; };   // [T*D][m][dim], [T*D][m], [T*D][m]

class PartitionedIndexService {
   public:
    PartitionedIndexService(GpuContext &gpu, const SystemConfig &cfg, GFunctions g) : gpu_(gpu), cfg_(cfg), g_(std::move(g)) {}

    void insert(int id, const std::vector<double> &vector) {                                                    // PIS:266-347
        if (frozen_) throw IllegalStateException("Index already finalized");
        if (id < 0) throw IllegalArgumentException("id / vector cannot be null");
        if (!ids_.empty() && (int)vector.size() != dim_) throw IllegalArgumentException("Mixed dimensions not supported in single index");
        dim_ = (int)vector.size();
        ids_.push_back(id);
        vecs_.insert(vecs_.end(), vector.begin(), vector.end());
    }
    // PIS:789-845: code every staged vector, GreedyPartitioner.build per (table, division), freeze -- on the device.
    void finalizeForSearch() {
        if (frozen_) return;
        const int64_t n = (int64_t)ids_.size();
        if (n < 1000) throw IllegalStateException("Cannot finalize index: only " + std::to_string(n) + " samples collected (< MIN_SAMPLE_SIZE)");
        const PaperConfig &pc = cfg_.paper;
        if (g_.dim != dim_ || (int64_t)g_.alpha.size() != (int64_t)pc.tables * pc.divisions * pc.m * dim_) throw IllegalStateException("GFunctionRegistry mismatch at finalize");
        gpu_.check(fspann_gfunctions_upload(gpu_.get(), dim_, pc.tables, pc.divisions, pc.m, pc.lambda, g_.alpha.data(), g_.r.data(), g_.omega.data()));
        std::vector<double> by_id((size_t)n * dim_);
        std::vector<int32_t> staged;                                    // insertion order: the 1000th and later first, then the first 999 (PIS:280-298, 821-831)
        for (int64_t i = 999; i < n; i++) staged.push_back(ids_[(size_t)i]);
        for (int64_t i = 0; i < 999; i++) staged.push_back(ids_[(size_t)i]);
        for (int64_t i = 0; i < n; i++) {
            if (ids_[(size_t)i] >= n) throw IllegalArgumentException("ids must be the ordinals 0..N-1 (FSA:501,515)");
            std::copy(vecs_.begin() + i * dim_, vecs_.begin() + (i + 1) * dim_, by_id.begin() + (int64_t)ids_[(size_t)i] * dim_);
        }
        gpu_.check(fspann_routing_build(gpu_.get(), n, by_id.data(), staged.data(), nullptr, nullptr, nullptr, nullptr));
        frozen_ = true;
        n_ = n;
        ids_.clear(); vecs_.clear();
    }
    bool isFrozen() const { return frozen_; }
    int numTables() const { return cfg_.paper.tables; }
    int dimension() const { return dim_; }
    int64_t size() const { return n_; }
    void setProbeOverride(int probes) { probeOverride_ = probes; }                                              // PIS:868-874
    void clearProbeOverride() { probeOverride_ = -1; }
    int getDefaultMaxProbes() const { return 5; }                                                               // PIS:93
    int effectiveMaxProbes() const { return probeOverride_ > 0 ? probeOverride_ : (cfg_.runtime.probeOverride > 0 ? cfg_.runtime.probeOverride : 5); }
    long long hardCap() const { return std::max<long long>(cfg_.runtime.maxGlobalCandidates, cfg_.runtime.refinementLimit); }     // PIS:612-615
    int W() const { return (cfg_.paper.m * cfg_.paper.lambda + 63) / 64; }

    // PIS:592-715: ordered (id, hammingScore) list, first `limit` entries
    std::vector<std::pair<int, int>> lookupCandidatesWithScores(const QueryToken &token, int limit = -1) {
        if (!frozen_) throw IllegalStateException("Index not finalized");
        if (token.bitCodes.empty()) throw IllegalStateException("MSANNP violation: QueryToken missing BitSet codes");
        if (token.numTables != cfg_.paper.tables) throw IllegalStateException("Token tables mismatch");
        if (token.dimension != dim_) return {};
        const int B = limit > 0 ? limit : cfg_.runtime.refinementLimit;
        std::vector<int32_t> ids((size_t)B, -1), sc((size_t)B, -1);
        int32_t n = 0, raw = 0, uq = 0;
        gpu_.check(fspann_route_batch(gpu_.get(), 1, token.bitCodes.data(), effectiveMaxProbes(), hardCap(), cfg_.runtime.hammingPrefilterThreshold, B,
                                      ids.data(), sc.data(), &n, &raw, &uq));
        lastRawVisited_ = raw;
        std::vector<std::pair<int, int>> out;
        for (int i = 0; i < n; i++) out.emplace_back(ids[(size_t)i], sc[(size_t)i]);
        return out;
    }
    int getLastRawCandidateCount() const { return lastRawVisited_; }
    GpuContext &gpu() { return gpu_; }
    const SystemConfig &config() const { return cfg_; }
    const GFunctions &registry() const { return g_; }

   private:
    GpuContext &gpu_;
    SystemConfig cfg_;
    GFunctions g_;
    bool frozen_ = false;
    int dim_ = 0, probeOverride_ = -1, lastRawVisited_ = 0;
    int64_t n_ = 0;
    std::vector<int32_t> ids_;
    std::vector<double> vecs_;
};

class QueryTokenFactory {
   public:
    using IvSource = std::function<std::vector<uint8_t>()>;         // EncryptionUtils.generateIV (SecureRandom); injectable for tests
    QueryTokenFactory(PartitionedIndexService &index, KeyManager &keys, IvSource iv) : index_(index), keys_(keys), iv_(std::move(iv)) {}

    QueryToken create(const std::vector<double> &vec, int topK) {                                               // QTF:63-167
        if (vec.empty()) throw IllegalArgumentException("query vector is null");
        if (topK <= 0) throw IllegalArgumentException("topK must be > 0");
        const GFunctions &g = index_.registry();
        const PaperConfig &pc = index_.config().paper;
        if (g.alpha.empty()) throw IllegalStateException("GFunctionRegistry not initialized. Build index first.");
        if (g.dim != (int)vec.size()) throw IllegalStateException("GFunctionRegistry mismatch: dimension=" + std::to_string(g.dim));
        QueryToken t;
        t.bitCodes.assign((size_t)pc.tables * pc.divisions * index_.W(), 0);
        index_.gpu().check(fspann_tokengen_batch(index_.gpu().get(), 1, vec.data(), t.bitCodes.data()));        // Coding.C per (t,d)
        t.version = keys_.getCurrentVersion();
        t.iv = iv_();
        t.encryptedQuery = detail::gcm_encrypt(keys_.getVersion(t.version), t.iv, detail::be_doubles(vec.data(), (int)vec.size()));   // QTF:152-154
        t.topK = topK; t.numTables = pc.tables; t.dimension = (int)vec.size(); t.lambda = pc.lambda;
        t.encryptionContext = "dim_" + std::to_string(t.dimension) + "_v" + std::to_string(t.version);
        return t;
    }

    // QTF:182-200: the same token (codes, IV, ciphertext, version) with another topK -- no new TokenGen, no new encryption
    QueryToken derive(const QueryToken *tok, int newTopK) const {
        if (!tok) throw IllegalArgumentException("token is null");
        if (newTopK <= 0) throw IllegalArgumentException("newTopK must be > 0");
        QueryToken t = *tok;
        t.topK = newTopK;
        return t;
    }

   private:
    PartitionedIndexService &index_;
    KeyManager &keys_;
    IvSource iv_;
};

class QueryServiceImpl {
   public:
    QueryServiceImpl(PartitionedIndexService &index, KeyManager &keys) : index_(index), keys_(keys) {}

    std::vector<QueryResult> search(const QueryToken *token) {                                                  // QSI:100-352
        if (!token) return {};                                                                                  // QSI:102
        return searchBatch({*token})[0];
    }
    std::vector<std::vector<QueryResult>> searchBatch(const std::vector<QueryToken> &tokens) {
        if (!index_.isFrozen()) throw IllegalStateException("Index not finalized");
        const int Q = (int)tokens.size();
        std::vector<std::vector<QueryResult>> out((size_t)Q);
        if (Q == 0) return out;
        const int dim = index_.dimension(), k = tokens[0].topK;
        const size_t code_words = (size_t)index_.config().paper.tables * index_.config().paper.divisions * index_.W();
        std::vector<double> qs;
        std::vector<uint64_t> codes;                                 // the tokens' OWN codes: Route runs on token.getBitCodes() (PIS:600)
        std::vector<int> keep;
        for (int i = 0; i < Q; i++) {
            const QueryToken &t = tokens[(size_t)i];
            if (t.topK != k) throw IllegalArgumentException("searchBatch needs one topK per batch (derive tokens per K like FSA:634)");
            if (t.bitCodes.empty()) throw IllegalStateException("MSANNP violation: QueryToken missing BitSet codes");          // PIS:604-606
            if (t.numTables != index_.config().paper.tables || t.bitCodes.size() != code_words) throw IllegalStateException("Token tables mismatch");
            std::vector<uint8_t> key;
            try { key = keys_.getVersion(t.version); } catch (const IllegalArgumentException &) { key = keys_.getVersion(keys_.getCurrentVersion()); }   // QSI:124-129
            const std::vector<double> q = detail::gcm_decrypt_doubles(key, t.iv, t.encryptedQuery);
            bool finite = (int)q.size() == dim;
            for (double v : q) finite = finite && std::isfinite(v);
            if (!finite) continue;                                                                              // QSI:137 -> empty result
            qs.insert(qs.end(), q.begin(), q.end());
            codes.insert(codes.end(), t.bitCodes.begin(), t.bitCodes.end());
            keep.push_back(i);
        }
        const int R = (int)keep.size();
        if (R == 0) return out;
        const RuntimeConfig &rt = index_.config().runtime;
        std::vector<int32_t> ids((size_t)R * k), nret((size_t)R);
        std::vector<double> dist((size_t)R * k);
        counters_.assign((size_t)R * FSPANN_COUNTERS, 0);
        const int rc = fspann_search_tokens(index_.gpu().get(), R, codes.data(), qs.data(), k, index_.effectiveMaxProbes(), index_.hardCap(),
                                            refineOverride_ > 0 ? refineOverride_ : rt.refinementLimit, rt.hammingPrefilterThreshold, ids.data(),
                                            dist.data(), nret.data(), counters_.data());
        index_.clearProbeOverride();                                                                            // QSI:342-346 (finally: also when the search fails)
        index_.gpu().check(rc);
        for (int j = 0; j < R; j++)
            for (int r = 0; r < nret[(size_t)j]; r++) out[(size_t)keep[(size_t)j]].push_back({std::to_string(ids[(size_t)j * k + r]), dist[(size_t)j * k + r]});
        const int64_t *c = &counters_[(size_t)(R - 1) * FSPANN_COUNTERS];                                       // getLast* describe the last query
        lastCandTotal_ = (int)c[0]; lastCandKept_ = (int)c[1]; lastCandDecrypted_ = (int)c[2]; lastReturned_ = (int)c[3];
        return out;
    }
    void setRefinementLimit(int limit) { refineOverride_ = limit; }
    void clearRefinementLimit() { refineOverride_ = -1; }
    int getLastCandTotal() const { return lastCandTotal_; }
    int getLastCandKept() const { return lastCandKept_; }
    int getLastCandDecrypted() const { return lastCandDecrypted_; }
    int getLastReturned() const { return lastReturned_; }
    // reencTracker.record(touched) (QSI:348-350): ids that reached verdict OK since the last call
    std::vector<int32_t> drainTouched() {
        const int64_t words = (index_.size() + 31) / 32;
        std::vector<uint32_t> bm((size_t)words);
        index_.gpu().check(fspann_touched_fetch(index_.gpu().get(), bm.data(), words, 1));
        std::vector<int32_t> out;
        for (int64_t i = 0; i < index_.size(); i++) if (bm[(size_t)(i >> 5)] >> (i & 31) & 1u) out.push_back((int32_t)i);
        return out;
    }

   private:
    PartitionedIndexService &index_;
    KeyManager &keys_;
    int refineOverride_ = -1, lastCandTotal_ = 0, lastCandKept_ = 0, lastCandDecrypted_ = 0, lastReturned_ = 0;
    std::vector<int64_t> counters_;
};

class ForwardSecureANNSystem {
   public:
    ForwardSecureANNSystem(const SystemConfig &cfg, int dim, std::vector<uint8_t> masterKey, GFunctions g, QueryTokenFactory::IvSource iv, int device = 0)
        : cfg_(cfg), dim_(dim), gpu_(device), keys_(std::move(masterKey)), index_(gpu_, cfg, std::move(g)), iv_(iv), tokenFactory_(index_, keys_, iv),
          queryService_(index_, keys_) {
        gpu_.check(fspann_keys_set(gpu_.get(), 1, keys_.derive(1).data()));
    }
    // FSA:479-560: id = ordinal; encryptToPoint (AGC:55-112) for the whole batch on the device; the host keeps the persistent mirror
    void batchInsert(const std::vector<double> &vectors) {
        if (vectors.size() % (size_t)dim_) throw IllegalArgumentException("Expected vector length " + std::to_string(dim_));
        const int64_t n = (int64_t)(vectors.size() / (size_t)dim_);
        std::vector<int32_t> ids((size_t)n);
        store_iv.resize((size_t)n * 12);
        for (int64_t i = 0; i < n; i++) {
            ids[(size_t)i] = (int32_t)i;
            const std::vector<uint8_t> iv = iv_();
            std::copy(iv.begin(), iv.end(), store_iv.begin() + i * 12);
            index_.insert((int)i, std::vector<double>(vectors.begin() + i * dim_, vectors.begin() + (i + 1) * dim_));
        }
        const int v = keys_.getCurrentVersion();
        store_ct.resize((size_t)n * (8 * (size_t)dim_ + 16));
        store_ver.assign((size_t)n, v);
        gpu_.check(fspann_encrypt_batch(gpu_.get(), n, dim_, ids.data(), vectors.data(), store_iv.data(), v, store_ct.data()));
        gpu_.check(fspann_store_upload(gpu_.get(), n, dim_, store_iv.data(), store_ct.data(), store_ver.data()));
    }
    void finalizeForSearch() { index_.finalizeForSearch(); }                                                    // FSA:977
    QueryToken createToken(const std::vector<double> &q, int topK, int dim) {                                   // FSA:1673-1698
        if (!index_.isFrozen()) throw IllegalStateException("Index is not finalized; call finalizeForSearch() before querying");
        if ((int)q.size() != dim || dim != dim_) throw IllegalArgumentException("Query dimension mismatch: expected=" + std::to_string(dim_));
        return tokenFactory_.create(q, topK);
    }
    int rotateKeyOnly() {                                                                                       // KRS:292-298
        const int v = keys_.rotateKey();
        gpu_.check(fspann_keys_set(gpu_.get(), v, keys_.derive(v).data()));
        return v;
    }
    // KRS:215-289 with the AES-GCM work on the device; returns ReencryptReport.reencrypted and refreshes the host mirror
    int64_t reencryptTouched(const std::vector<int32_t> &ids, int targetVersion) {
        const int64_t n = (int64_t)ids.size();
        if (n == 0) return 0;
        std::vector<uint8_t> fresh((size_t)n * 12), done((size_t)n), iv((size_t)n * 12), ct((size_t)n * (8 * (size_t)dim_ + 16));
        for (int64_t i = 0; i < n; i++) { const std::vector<uint8_t> x = iv_(); std::copy(x.begin(), x.end(), fresh.begin() + i * 12); }
        int64_t cnt = 0;
        gpu_.check(fspann_migrate(gpu_.get(), n, ids.data(), fresh.data(), targetVersion, done.data(), iv.data(), ct.data(), &cnt));
        const size_t row = 8 * (size_t)dim_ + 16;
        for (int64_t i = 0; i < n; i++)
            if (done[(size_t)i]) {                                                                              // metadataManager.saveEncryptedPoint (KRS:268)
                std::copy(iv.begin() + i * 12, iv.begin() + (i + 1) * 12, store_iv.begin() + (int64_t)ids[(size_t)i] * 12);
                std::copy(ct.begin() + i * (int64_t)row, ct.begin() + (i + 1) * (int64_t)row, store_ct.begin() + (int64_t)ids[(size_t)i] * (int64_t)row);
                store_ver[(size_t)ids[(size_t)i]] = targetVersion;
            }
        return cnt;
    }
    bool retire(int version) {                                                                                  // KM:287-294: refused while records are bound
        if (std::find(store_ver.begin(), store_ver.end(), version) != store_ver.end()) return false;
        keys_.retire(version);
        gpu_.check(fspann_keys_retire(gpu_.get(), version));
        return true;
    }
    PartitionedIndexService &index() { return index_; }
    QueryServiceImpl &queryService() { return queryService_; }
    QueryTokenFactory &tokenFactory() { return tokenFactory_; }
    KeyManager &keys() { return keys_; }
    GpuContext &gpu() { return gpu_; }
    std::vector<uint8_t> store_iv, store_ct;   // host mirror of the encrypted store (what RocksDB + .point files hold)
    std::vector<int32_t> store_ver;

   private:
    SystemConfig cfg_;
    int dim_;
    GpuContext gpu_;
    KeyManager keys_;
    PartitionedIndexService index_;
    QueryTokenFactory::IvSource iv_;
    QueryTokenFactory tokenFactory_;
    QueryServiceImpl queryService_;
};

}  // namespace fspann
