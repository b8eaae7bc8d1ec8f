/*
 * fspann_gpu.h -- C ABI of libfspann_gpu.so: the B200 (sm_100a) implementation of FSPANN's query hot path
 * (TokenGen -> Route -> Refine).
 *
 * The reference (pure Java, /root/reference/fsp-anns-parent) has no FFI for this path; its seam is
 *   QueryService.search(QueryToken)                       query/src/main/java/com/fspann/query/service/QueryService.java:10-12
 *   QueryTokenFactory.create(double[], int)               query/src/main/java/com/fspann/query/core/QueryTokenFactory.java:63
 *   PartitionedIndexService.lookupCandidatesWithScores    index/src/main/java/com/fspann/index/paper/PartitionedIndexService.java:592
 *   PartitionedIndexService.loadPointIfActive             .../PartitionedIndexService.java:717
 *   CryptoService.decryptFromPoint                        crypto/src/main/java/com/fspann/crypto/AesGcmCryptoService.java:126
 *   KeyLifeCycleService.getVersion                        keymanagement/src/main/java/com/fspann/key/KeyRotationServiceImpl.java:82
 * Each entry point below names the Java member it replaces.  A JNI / Panama binding that a maintainer would
 * add on the Java side is shown in INTEGRATION.md.
 *
 * Conventions (mirroring the reference's):
 *  - plain C, no torch / CUDA types in signatures; every array is caller-owned and copied (the reference clones
 *    every byte[] at its getters, common/.../QueryToken.java:76-93), so no host pointer is retained;
 *  - one context per GPU; calls on one context are serialised on its stream; distinct contexts are independent;
 *  - return 0 on success, negative on error; fspann_last_error(ctx) gives the text:
 *        FSPANN_E_ARG    <-> IllegalArgumentException / NullPointerException (bad argument, dimension mismatch)
 *        FSPANN_E_STATE  <-> IllegalStateException (index not finalized / registry mismatch, PIS:594, QTF:67-88)
 *        FSPANN_E_CUDA   <-> a CUDA runtime failure (there is NO CPU fallback)
 *        FSPANN_E_NOMEM  <-> device allocation failure
 *  - per-candidate failures never fail a call: they are reported as verdict codes and counted (QSI:242-270);
 *  - ids are the non-negative 32-bit integers whose decimal strings the reference's facade uses as point ids
 *    (api/.../ForwardSecureANNSystem.java:501,515).  Routing order depends on java.lang.String.hashCode of that
 *    string, which the kernels recompute.
 *  - "_dev" variants take DEVICE pointers (inputs already resident in HBM) and enqueue on the context stream
 *    without synchronising; everything else takes HOST pointers and returns after the results are on the host.
 */
#ifndef FSPANN_GPU_H
#define FSPANN_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSPANN_OK 0
#define FSPANN_E_ARG (-1)
#define FSPANN_E_STATE (-2)
#define FSPANN_E_CUDA (-3)
#define FSPANN_E_NOMEM (-4)

/* verdict codes per refined candidate (QSI:242-270 counters notFound / decryptError / invalidVector) */
#define FSPANN_V_OK 0         /* decrypted, tag verified, finite: distance evaluated, id is "touched" (QSI:262) */
#define FSPANN_V_NOT_FOUND 1  /* loadPointIfActive returned null: deleted / unknown id (PIS:717-724)           */
#define FSPANN_V_NO_KEY 2     /* keyService.getVersion threw: version unknown or retired (KRS:82-88)           */
#define FSPANN_V_TAG_FAIL 3   /* AES-GCM tag mismatch (AGC:159-165)                                            */
#define FSPANN_V_NON_FINITE 4 /* plaintext holds NaN/Inf (QSI:253, QSI:407-413)                                */
#define FSPANN_V_OTHER_SHARD 0xFD /* sharded store only: the record is held by another GPU                       */

#define FSPANN_BLOCK 64       /* DEFAULT_GREEDY_BLOCK_SIZE (PIS:92) */
#define FSPANN_MAX_KEYS 16    /* live key versions held on the device */
#define FSPANN_COUNTERS 6     /* per query: candTotal(raw), candKept(unique), candDecrypted, returned, retried, refined */

typedef struct fspann_ctx fspann_ctx;

/* ---- context ------------------------------------------------------------------------------------------ */
int fspann_ctx_create(int device, fspann_ctx **out);
void fspann_ctx_destroy(fspann_ctx *ctx);
const char *fspann_last_error(const fspann_ctx *ctx);
/* cudaStream_t of the context (as void*), so a harness can record CUDA events around "_dev" calls. */
void *fspann_ctx_stream(fspann_ctx *ctx);
int fspann_ctx_sync(fspann_ctx *ctx);
/* number of kernels this library has launched on the context since creation (bench.py's gpu_launches). */
int64_t fspann_ctx_launch_count(const fspann_ctx *ctx);

/* Tuning / test switches: "route_general" = 1 forces the general Route kernel (sequential groups, exact HARD_CAP
 * semantics) even where the shared-memory fast path applies; "route_v1" = 1 keeps the fast path on its one-CTA-per-SM kernel (the
 * default for refinementLimit <= 1024 is the two-CTA kernel, which hands the queries it cannot hold to the one-CTA kernel);
 * "route_small_v1" = 0 keeps batches of at most one query per SM on the two-CTA kernel too (by default they take the one-CTA kernel: lower
 * latency, no overflow pass); "route_wl_extra" = n >= 0 clamps the fast path's dedicated worklist to n entries (test hook for the overflow fallbacks),
 * -1 = automatic; "h2d_overlap" = n (default 2, 0 = off): the host-pointer search entries upload a batch of >= 4096 queries in n
 * chunks on a copy stream so TokenGen + Route of a chunk overlap the PCIe copy of the next (with supplied codes: the whole query
 * upload overlaps Route); "tokengen_exact" = 1 runs the exact FP64 TokenGen
 * kernel alone instead of a pre-filter + exact re-check; "tokengen_mode" = 0 (default) runs the pre-filter's contraction on the tensor
 * cores (tcgen05.mma on BF16-split operands, accumulator in TMEM) when the shape allows it, 2 = on the FP32 FMA pipe (the codes are
 * identical in every mode); "graphs" = 1 (default) lets a search of <= 64 queries replay its first pass (TokenGen .. counters and the
 * retry decision) as ONE captured CUDA graph from the third call with the same shape and buffers on (the second call captures; any upload,
 * key, option or buffer change drops the match), 0 = always launch kernel by kernel -- results are identical either way.  fspann_get_info:
 * "graph_captures" / "graph_replays" (how many such graphs were built / launched on this context), "last_tokengen_path" (1 exact, 2 FP32 pre-filter, 3 tensor-core pre-filter),
 * "last_route_path" (1 fast, 2 general), "last_route_v2" (1: the two-CTA fast kernel ran), "route_overflowed" (queries of the last call it
 * handed to the one-CTA kernel), (general path only: 1 if a bin of a query's bestScore map reached 9 entries in the
 * last call -- the JDK treeifies such a bin and its iteration order is not modelled), "sm_count", "build_treeified", "tokengen_rechecked" / "tokengen_overflow" (projections
 * the last TokenGen launch had to re-check exactly / whether its list overflowed and the exact kernel recomputed the batch). */
int fspann_set_option(fspann_ctx *ctx, const char *name, int64_t value);
int64_t fspann_get_info(fspann_ctx *ctx, const char *name);

/* ---- Setup state crossing the boundary ----------------------------------------------------------------- */
/* Routing state I = GFunctions + frozen partitions: GFunctionRegistry contents (GFR:63-147; Coding.GFunction
 * alpha/r/omega, Coding:52-97) and every DivisionState.partitions after finalizeForSearch (PIS:789-845, GP:13-32).
 *   alpha[T*D][m][dim], r[T*D][m], omega[T*D][m] FP64;   n_ids = points per (table,division);
 *   P = ceil(n_ids/64) partitions per (t,d);  min_key/max_key int64 [T*D][P];  rep_code uint64 [T*D][P][W],
 *   W = ceil(m*lambda/64), bit p of a code = BitSet bit p (word p>>6, bit p&63);  ids int32 [T*D][n_ids] in
 *   partition order (partition i owns slots [64i, min(64i+64, n_ids))).
 * Immutable until the next upload (the reference freezes the index, PIS:842). */
int fspann_routing_upload(fspann_ctx *ctx, int32_t dim, int32_t T, int32_t D, int32_t m, int32_t lambda,
                          const double *alpha, const double *r, const double *omega, int64_t n_ids,
                          const int64_t *min_key, const int64_t *max_key, const uint64_t *rep_code,
                          const int32_t *ids);

/* GFunctions only: lets TokenGen run before any partition exists (Setup-side bulk coding of the base set,
 * PIS:331-346).  Route / search on such a context fail with FSPANN_E_STATE ("Index not finalized"). */
int fspann_gfunctions_upload(fspann_ctx *ctx, int32_t dim, int32_t T, int32_t D, int32_t m, int32_t lambda,
                             const double *alpha, const double *r, const double *omega);

/* Setup-side index build on the device.  Needs the GFunctions (fspann_gfunctions_upload).  Codes every vector for
 * every (table, division) with the TokenGen kernel (the coding loop of PartitionedIndexService.insert / finalizeForSearch,
 * PIS:331-346, 821-831) and runs GreedyPartitioner.build (GP:37-76) per division exactly as finalizeForSearch does: entries in
 * java.util.HashMap<String,BitSet>(N) iteration order (PIS:413-420), stable sort by computeKey (GP:51, 87-96), blocks of 64,
 * repCode = code of entry i + ((end-i-1)>>>1) (GP:60).  The result is installed as the routing state (= fspann_routing_upload).
 *   vectors FP64 [N][dim] indexed by id (ids are the ordinals 0..N-1, FSA:501,515); staged_ids int32 [N] = insertion order into
 *   the staging map (the 1000th and later vectors first, then the first 999: PIS:280-298, 821-831).
 *   Optional host outputs (NULL to skip), fspann_routing_upload's layout: min_key/max_key int64 [T*D][P], rep_code uint64
 *   [T*D][P][W], ids int32 [T*D][N].  fspann_get_info("build_treeified") = 1 if a HashMap bin held >= 9 entries (the JDK
 *   treeifies it and its iteration order is unspecified; does not happen for decimal ids at the BASELINE sizes).
 * Errors: FSPANN_E_STATE without GFunctions or with N < 1000 (PIS:803-819); FSPANN_E_ARG for NaN/Inf or ids that are not 0..N-1. */
int fspann_routing_build(fspann_ctx *ctx, int64_t N, const double *vectors, const int32_t *staged_ids, int64_t *min_key_out,
                         int64_t *max_key_out, uint64_t *rep_code_out, int32_t *ids_out);

/* The same build in pieces, for a base set that does not fit the host at once (BASELINE config 4: 100 M x 96): begin(N) reserves the code
 * array [N][T*D][W] in HBM, add / add_dev code the vectors of ids first_id .. first_id+n-1 (host FP64 rows, or rows already resident in
 * HBM) in any order, each id exactly once, and finish runs GreedyPartitioner.build per division and installs the routing state.
 * staged_ids == NULL means the facade's own insertion order (ids 999 .. N-1, then 0 .. 998: PIS:280-298, 821-831), generated on the
 * device.  fspann_routing_build = begin + add(0, N) + finish. */
int fspann_routing_build_begin(fspann_ctx *ctx, int64_t N);
int fspann_routing_build_add(fspann_ctx *ctx, int64_t first_id, int64_t n, const double *vectors);
int fspann_routing_build_add_dev(fspann_ctx *ctx, int64_t first_id, int64_t n, const double *d_vectors);
int fspann_routing_build_finish(fspann_ctx *ctx, const int32_t *staged_ids, int64_t *min_key_out, int64_t *max_key_out,
                                uint64_t *rep_code_out, int32_t *ids_out);

/* metadata.isDeleted(id) (common/.../RocksDBMetadataManager.java:203-224): flags[id] != 0 => deleted.
 * n may be 0 / flags NULL to clear.  Consulted by Route (PIS:739) and Refine (PIS:718). */
int fspann_deleted_set(fspann_ctx *ctx, const uint8_t *flags, int64_t n);

/* Record store Store_t: EncryptedPoint fields (EP:18-26) for ids 0..N-1, packed:
 *   iv uint8 [N][12], ct uint8 [N][8*dim+16] (ciphertext || 128-bit tag, Java doFinal layout),
 *   key_version int32 [N].  Replaces RocksDBMetadataManager.loadEncryptedPoint (RDB:530-544) on the hot path. */
int fspann_store_upload(fspann_ctx *ctx, int64_t N, int32_t dim, const uint8_t *iv, const uint8_t *ct,
                        const int32_t *key_version);
/* Database-sharded deployment (BASELINE config 4): this context holds only the records with global ids
 * [id_base, id_base + N) of a store whose ids span [0, n_global).  AAD and candidate ids stay GLOBAL; a candidate that
 * lives in another shard gets verdict 0xFD here and is refined by its owner.  fspann_store_upload = shard (0, N, N). */
int fspann_store_upload_shard(fspann_ctx *ctx, int64_t id_base, int64_t N, int64_t n_global, int32_t dim,
                              const uint8_t *iv, const uint8_t *ct, const int32_t *key_version);
/* Setup of a shard that is produced on the device (config 4): alloc_shard reserves an empty shard for the global ids
 * [id_base, id_base + N); fspann_store_encrypt_dev runs encryptToPoint (AGC:55-112; AAD id:<id>|v:<version>|d:<dim>, EP:80-83) for the
 * device-resident FP64 vectors [n][dim] of ids first_id .. first_id+n-1 with the device-resident IVs [n][12] and writes the records
 * straight into the HBM store.  A row never written keeps key version 0 and verdict NO_KEY. */
int fspann_store_alloc_shard(fspann_ctx *ctx, int64_t id_base, int64_t N, int64_t n_global, int32_t dim);
int fspann_store_encrypt_dev(fspann_ctx *ctx, int64_t first_id, int64_t n, const double *d_vectors, const uint8_t *d_ivs, int32_t version);
/* In-place replacement of n records: the result of Migrate = reencryptTouched (KRS:250-266).  Atomic with
 * respect to later batches (stream ordered). */
int fspann_store_update(fspann_ctx *ctx, int64_t n, const int32_t *ids, const uint8_t *iv, const uint8_t *ct,
                        const int32_t *key_version);

/* ---- Migrate on the device: KeyRotationServiceImpl.reencryptTouched
 * (keymanagement/src/main/java/com/fspann/key/KeyRotationServiceImpl.java:215-289).  For every id of the list, in order and
 * once (new LinkedHashSet<>(touchedIds), KRS:232): skip if the store does not hold it (null load, KRS:243), if its stored
 * version is >= target_version (KRS:248), if its key is gone or its tag fails ("forward-secure skip", KRS:277-279); otherwise
 * AES-GCM decrypt under the stored version and re-encrypt under target_version with AAD id|v:target|d and the fresh IV
 * fresh_ivs[i] (the 12 bytes cryptoService.encrypt would draw from SecureRandom, AGC:69-71), IN PLACE in the HBM store, so the
 * next batch already sees the migrated record.  The plaintext never leaves registers.
 *   reencrypted_out uint8 [n] (may be NULL): 1 where list entry i was re-encrypted;
 *   iv_out uint8 [n][12], ct_out uint8 [n][8*dim+16] (may be NULL): the new record for the host's RocksDB / .point file
 *   (metadataManager.saveEncryptedPoint, KRS:268), valid where reencrypted_out[i] = 1;
 *   *n_reencrypted_out = ReencryptReport.reencrypted.  target_version must be a live key (fspann_keys_set). */
int fspann_migrate(fspann_ctx *ctx, int64_t n, const int32_t *ids, const uint8_t *fresh_ivs, int32_t target_version,
                   uint8_t *reencrypted_out, uint8_t *iv_out, uint8_t *ct_out, int64_t *n_reencrypted_out);

/* ---- Setup-side bulk encryption: AesGcmCryptoService.encryptToPoint
 * (crypto/src/main/java/com/fspann/crypto/AesGcmCryptoService.java:55-112) for n vectors: big-endian FP64 serialisation
 * (AGC:240-259), AES-256-GCM under key `version` with the caller's 12-byte IVs and AAD "id:<id>|v:<version>|d:<dim>"
 * (EP:80-83).  vectors FP64 [n][dim] -> ct_out uint8 [n][8*dim+16] (ciphertext || tag, Java doFinal layout), ready for
 * fspann_store_upload.  Does not touch the store; the device copy of the vectors is wiped before returning. */
int fspann_encrypt_batch(fspann_ctx *ctx, int64_t n, int32_t dim, const int32_t *ids, const double *vectors, const uint8_t *ivs,
                         int32_t version, uint8_t *ct_out);

/* Key ring K_t: KeyLifeCycleService.getVersion(v) -> 32-byte AES key (KRS:82-88, KM:221-237).
 * set = Rotate made version v available; retire = Retire deleted it (KM:274-317): records still bound to a
 * retired version then yield FSPANN_V_NO_KEY, exactly as the Java path counts a decryptError. */
int fspann_keys_set(fspann_ctx *ctx, int32_t version, const uint8_t key[32]);
int fspann_keys_retire(fspann_ctx *ctx, int32_t version);

/* ---- TokenGen (a2-a3): Coding.C for every (table, division) of every query (QTF:98-131, Coding:250-301) ---
 * queries FP64 [Q][dim] -> codes uint64 [Q][T*D][W].  A query holding NaN/Inf fails the whole call with
 * FSPANN_E_ARG (Coding:355-360 requireVector). */
int fspann_tokengen_batch(fspann_ctx *ctx, int64_t Q, const double *queries, uint64_t *codes_out);
/* Device-resident variant (no copies, no synchronisation, no NaN/Inf scan): Setup-side bulk coding of vectors already in HBM. */
int fspann_tokengen_batch_dev(fspann_ctx *ctx, int64_t Q, const double *d_queries, uint64_t *d_codes);

/* ---- Route (a6-a12): lookupCandidatesWithScores + the first-B cut of QueryServiceImpl stage A.5 ------------
 * (PIS:592-753, QSI:153-214).  probes = effectiveMaxProbes() (PIS:880-888), hard_cap = max(maxGlobalCandidates,
 * refinementLimit) (PIS:612-615), B = effective refinementLimit (> 0), ham_threshold = hammingPrefilterThreshold.
 * Outputs per query q: cand_ids/cand_scores [Q][B] (first n_cand[q] valid, in the reference's order),
 * raw_seen[q] = getLastRawCandidateCount(), unique[q] = number of distinct candidates (lastCandKept). */
int fspann_route_batch(fspann_ctx *ctx, int64_t Q, const uint64_t *codes, int32_t probes, int64_t hard_cap,
                       int32_t ham_threshold, int32_t B, int32_t *cand_ids_out, int32_t *cand_scores_out,
                       int32_t *n_cand_out, int32_t *raw_seen_out, int32_t *unique_out);

/* ---- Refine (a13-a17): stage B + C of QueryServiceImpl.search (QSI:238-322) ---------------------------------
 * For each query: candidates cand_ids[q][0..n_cand[q]) in order -> load, key lookup, AES-256-GCM verify+decrypt
 * under the record's stored key version, finite check, exact sequential FP64 L2, stable top-k.
 * cand_stride = row stride of cand_ids / verdict_out (>= max n_cand).  Outputs: topk_ids/topk_dist [Q][k]
 * (first n_ret[q] valid), verdict_out uint8 [Q][cand_stride] (may be NULL), n_decrypted_out[q] (may be NULL).
 * Plaintext vectors exist only in registers / shared memory and are never written to global memory. */
int fspann_refine_batch(fspann_ctx *ctx, int64_t Q, const double *queries, const int32_t *cand_ids,
                        const int32_t *n_cand, int32_t cand_stride, int32_t k, int32_t *topk_ids_out,
                        double *topk_dist_out, int32_t *n_ret_out, uint8_t *verdict_out, int32_t *n_decrypted_out);

/* Same, plus topk_rank_out [Q][k]: the position of each result in its candidate list.  Per-shard results are merged
 * across GPUs on (distance, rank), which reproduces the reference's stable sort (QSI:298) exactly. */
int fspann_refine_batch_ex(fspann_ctx *ctx, int64_t Q, const double *queries, const int32_t *cand_ids,
                           const int32_t *n_cand, int32_t cand_stride, int32_t k, int32_t *topk_ids_out,
                           double *topk_dist_out, int32_t *topk_rank_out, int32_t *n_ret_out, uint8_t *verdict_out,
                           int32_t *n_decrypted_out);

/* ---- search (a4-a18): createToken + QueryServiceImpl.search for a batch, incl. the adaptive retry (QSI:327-337):
 * a query whose first pass returned < k results or decrypted < 10*k candidates is re-run once with 10 probes and
 * the second result is returned.  counters int64 [Q][FSPANN_COUNTERS] (may be NULL).  The retry decision is taken on the device; the
 * host reads back two integers with the results, so a batch without retries costs one stream synchronisation.  A query holding NaN/Inf
 * fails the call with FSPANN_E_ARG like createToken throws (Coding:357-359); the output buffers are then unspecified and the
 * offending queries have touched no record. */
int fspann_search_batch(fspann_ctx *ctx, int64_t Q, const double *queries, int32_t k, int32_t probes,
                        int64_t hard_cap, int32_t B, int32_t ham_threshold, int32_t *topk_ids_out,
                        double *topk_dist_out, int32_t *n_ret_out, int64_t *counters_out);

/* Device-resident variant: d_queries, d_topk_ids, d_topk_dist, d_n_ret, d_counters are DEVICE pointers; work is
 * enqueued on the context stream and the call returns without synchronising.  The adaptive retry needs the number of
 * retrying queries on the host (two integers, decided on the device), so it is applied only when allow_retry != 0 (which synchronises
 * once).  With allow_retry == 0 a query holding NaN/Inf returns empty (QSI:137) instead of failing the call. */
int fspann_search_batch_dev(fspann_ctx *ctx, int64_t Q, const double *d_queries, int32_t k, int32_t probes,
                            int64_t hard_cap, int32_t B, int32_t ham_threshold, int32_t allow_retry,
                            int32_t *d_topk_ids, double *d_topk_dist, int32_t *d_n_ret, int64_t *d_counters);

/* ---- search on the token's own codes: QueryServiceImpl.search(QueryToken) as written (QSI:100-352).  The reference routes on
 * token.getBitCodes() (index/.../PartitionedIndexService.java:600) -- the codes the CLIENT computed in QueryTokenFactory.create /
 * derive (query/.../QueryTokenFactory.java:98-131, 182-200) -- and never recomputes them from the decrypted query, so TokenGen runs
 * once per query (in fspann_tokengen_batch, when the token is made) and not again here.  codes uint64 [Q][T*D][W] as produced by
 * fspann_tokengen_batch; queries FP64 [Q][dim] = the decrypted token payloads (AGC:189-204).  Differences from fspann_search_batch:
 * no TokenGen kernel; a query holding NaN/Inf does not fail the call, it returns an empty list and touches nothing (QSI:137) while the
 * rest of the batch is served; codes == NULL fails with FSPANN_E_STATE ("QueryToken missing BitSet codes", PIS:604-606).  The adaptive
 * retry re-routes the same codes with 10 probes (QSI:327-337).  The _dev variant takes device pointers (see fspann_search_batch_dev). */
int fspann_search_tokens(fspann_ctx *ctx, int64_t Q, const uint64_t *codes, const double *queries, int32_t k, int32_t probes,
                         int64_t hard_cap, int32_t B, int32_t ham_threshold, int32_t *topk_ids_out, double *topk_dist_out,
                         int32_t *n_ret_out, int64_t *counters_out);
int fspann_search_tokens_dev(fspann_ctx *ctx, int64_t Q, const uint64_t *d_codes, const double *d_queries, int32_t k, int32_t probes,
                             int64_t hard_cap, int32_t B, int32_t ham_threshold, int32_t allow_retry, int32_t *d_topk_ids,
                             double *d_topk_dist, int32_t *d_n_ret, int64_t *d_counters);

/* ---- evaluation on the box ------------------------------------------------------------------------------
 * Exact ground truth = GroundtruthPrecompute.run (api/src/main/java/com/fspann/api/GroundtruthPrecompute.java:218-276): for every
 * query the K nearest base vectors by squared L2 computed like VecReader.l2sq (:144-163) -- per dimension a FLOAT subtraction
 * q[i] - b[i] widened to double, squares summed sequentially in FP64 -- ordered by (distance, id) ascending (:168-189).
 * base float32 [N][dim], queries float32 [Q][dim] (.fvecs content; .bvecs bytes are passed as the floats 0..255, which the
 * reference's float - int arithmetic equals); K in [1, min(N, 1024)] (callers clamp like kFinal, :238).
 * gt_ids_out int32 [Q][K]; gt_d2_out FP64 [Q][K] squared distances (may be NULL). */
int fspann_groundtruth(fspann_ctx *ctx, int64_t N, int32_t dim, const float *base, int64_t Q, const float *queries, int32_t K,
                       int32_t *gt_ids_out, double *gt_d2_out);
/* recall@K of ForwardSecureANNSystem.computeMetricsAtK (api/.../ForwardSecureANNSystem.java:785-794):
 * recall[q] = |{ i < min(K, n_ret[q]) : result_ids[q][i] in gt_ids[q][0..K) }| / K.  gt_stride >= K; n_ret may be NULL (= K). */
int fspann_recall_batch(fspann_ctx *ctx, int64_t Q, int32_t K, const int32_t *gt_ids, int32_t gt_stride, const int32_t *result_ids,
                        int32_t result_stride, const int32_t *n_ret, double *recall_out);

/* ---- database-sharded deployment (BASELINE config 4), device-resident building blocks; all pointers are DEVICE pointers and
 * the work is enqueued on the context stream without synchronising ------------------------------------------------------------
 * Route is query-parallel on the replicated routing index (each GPU routes its slice of the batch), Refine is data-parallel on
 * the sharded store (each GPU refines, for all queries, the candidates it holds), connected by two all-gathers (NCCL):
 *   fspann_route_batch_dev   = createToken's coding + lookupCandidatesWithScores + first-B cut (QTF:98-131, PIS:592-715, QSI:153-214)
 *   fspann_refine_batch_dev  = QSI:238-322 on this shard; d_topk_rank [Q][k] = position of each result in its candidate list
 *   fspann_merge_topk_dev    = the global stable top-k (QSI:298-316) from n_shards per-shard lists laid out [n_shards][Q][k]
 *                              (id = -1 pads), ordered by (distance, candidate rank). */
int fspann_route_batch_dev(fspann_ctx *ctx, int64_t Q, const double *d_queries, int32_t probes, int64_t hard_cap, int32_t B,
                           int32_t *d_cand_ids, int32_t *d_n_cand, int32_t *d_raw_seen, int32_t *d_unique);
int fspann_refine_batch_dev(fspann_ctx *ctx, int64_t Q, const double *d_queries, const int32_t *d_cand_ids, const int32_t *d_n_cand,
                            int32_t cand_stride, int32_t k, int32_t *d_topk_ids, double *d_topk_dist, int32_t *d_topk_rank,
                            int32_t *d_n_ret, int32_t *d_n_decrypted);
int fspann_merge_topk_dev(fspann_ctx *ctx, int32_t n_shards, int64_t Q, int32_t k, const double *d_dist, const int32_t *d_rank,
                          const int32_t *d_ids, int32_t *d_out_ids, double *d_out_dist, int32_t *d_out_n_ret);

/* ---- database-sharded search as ONE call (config 4): the NCCL collectives run inside this library -------------------------------
 * One context per GPU, each holding the replicated routing state and ITS shard of the store (fspann_store_upload_shard); the W contexts
 * form a communicator.  They may live in W processes (one per GPU) or in W threads of one process (one JVM, a thread per GPU).
 *   fspann_comm_unique_id : any one participant draws an id (ncclGetUniqueId) and hands the 128 bytes to the others;
 *   fspann_comm_init      : COLLECTIVE -- every participant calls it with the same id, n_ranks and its own rank (ncclCommInitRank on the
 *                           context's device).  n_ranks == 1 needs no id and no NCCL.  rank r must hold the r-th id range of the store.
 *   fspann_sharded_search_batch(_dev) : COLLECTIVE -- every participant calls it with the SAME query batch and parameters and receives
 *                           the same result, bit-identical to fspann_search_batch on an unsharded store: rank r codes + routes rows
 *                           [r*ceil(Q/W), ...) of the batch (QTF:98-131, PIS:592-715), the ordered candidate lists are all-gathered
 *                           (Q*B*4 bytes), every rank refines the candidates its shard holds for all queries (QSI:238-322), the per-shard
 *                           top-k (distance, candidate rank, id) is all-gathered and merged on (distance, rank) -- the reference's stable
 *                           sort (QSI:298) -- and the adaptive retry (QSI:327-337) is decided from the merged counts.  counters as in
 *                           fspann_search_batch (candDecrypted = sum over the shards).  The host-pointer variant rejects NaN/Inf queries
 *                           with FSPANN_E_ARG before any collective; the _dev variant takes device pointers and synchronises only for
 *                           the retry decision (allow_retry != 0).
 * NCCL (libnccl.so.2) is loaded at run time: the copy already present in the process, else the system's, else $FSPANN_NCCL_LIB.
 * Without it fspann_comm_unique_id / fspann_comm_init (n_ranks > 1) fail with FSPANN_E_STATE and everything else works. */
#define FSPANN_COMM_ID_BYTES 128
int fspann_comm_unique_id(uint8_t id_out[FSPANN_COMM_ID_BYTES]);
int fspann_comm_init(fspann_ctx *ctx, int32_t n_ranks, int32_t rank, const uint8_t id[FSPANN_COMM_ID_BYTES]);
int fspann_comm_destroy(fspann_ctx *ctx);
int fspann_sharded_search_batch(fspann_ctx *ctx, int64_t Q, const double *queries, int32_t k, int32_t probes, int64_t hard_cap, int32_t B,
                                int32_t *topk_ids_out, double *topk_dist_out, int32_t *n_ret_out, int64_t *counters_out);
int fspann_sharded_search_batch_dev(fspann_ctx *ctx, int64_t Q, const double *d_queries, int32_t k, int32_t probes, int64_t hard_cap,
                                    int32_t B, int32_t allow_retry, int32_t *d_topk_ids, double *d_topk_dist, int32_t *d_n_ret,
                                    int64_t *d_counters);
/* Device time of the last sharded call's first pass (CUDA events on the context stream), milliseconds: out[0] = TokenGen + Route of this
 * rank's slice, out[1] = all-gather of the candidate lists, out[2] = Refine on this shard (fspann_last_stage_ms splits it further),
 * out[3] = all-gather of the per-shard top-k + merge.  *gather_bytes_out = bytes this rank received through the collectives. */
int64_t fspann_sharded_last_stage_ms(fspann_ctx *ctx, float out[4], int64_t *gather_bytes_out);

/* Touched set (QSI:262, QSI:348-350 reencTracker.record): bitmap over ids (bit id&31 of word id>>5, N bits) of
 * every record that reached verdict OK since the last clear.  Feeds the host's selective re-encryption. */
int fspann_touched_fetch(fspann_ctx *ctx, uint32_t *bitmap_out, int64_t n_words, int32_t clear);

/* Per-stage device time in milliseconds of the last search/refine call (CUDA events on the context stream):
 * out[0]=tokengen, out[1]=route, out[2]=refine-group(count/scan/fill), out[3]=refine-verify (GHASH + tag),
 * out[4]=refine-decrypt+distance, out[5]=top-k.  Returns the number of kernel launches of that call. */
int64_t fspann_last_stage_ms(fspann_ctx *ctx, float out[6]);

/* ---- debug build only (compiled with -DFSPANN_DEBUG_TAP): decrypted plaintext for parity tests -------------
 * Returns FSPANN_E_STATE in production builds, where plaintext never reaches global memory. */
int fspann_debug_decrypt(fspann_ctx *ctx, int64_t n, const int32_t *ids, double *plaintext_out, uint8_t *verdict_out);

#ifdef __cplusplus
}
#endif
#endif /* FSPANN_GPU_H */
