// Source only: this repository's image has no JDK, so this file is not compiled here.  A maintainer of Mehran-Memon/fspann-query-system
// compiles it against the reference's own modules (api, index, query, crypto, keymanagement, common, config) and runs it ONCE to pin the
// parity of the CUDA path against the real Java implementation:
//
//     mvn -q -pl it -am dependency:build-classpath -Dmdep.outputFile=cp.txt
//     javac -cp "$(cat cp.txt):api/target/classes:..." -d out integration/java/com/fspann/gpu/GoldenDump.java
//     java  -cp "out:$(cat cp.txt):..." com.fspann.gpu.GoldenDump <repo>/tests/golden/reference
//
// It drives the STOCK reference (ForwardSecureANNSystem, constructed exactly like it/src/test/java/com/fspann/it/BaseUnifiedIT.java:36-113)
// over a small deterministic SIFT-shaped data set and writes, as little-endian raw arrays plus manifest.json, every intermediate the CUDA
// path claims to reproduce bit for bit:
//   inputs      base / queries (FP64), the GFunction arrays of the registry (alpha, r, omega: they cross the C ABI as data, SURVEY 8a1)
//   store       iv, ciphertext||tag and key version of every EncryptedPoint, the session key of every version
//   TokenGen    token.getBitCodes() as BitSet.toLongArray words                       (QueryTokenFactory.create, QTF:63-167)
//   Route       lookupCandidatesWithScores: ordered ids + Hamming scores, raw count   (PIS:592-715)
//   Refine      QueryServiceImpl.search: ids + FP64 distances, the getLast* counters  (QSI:100-352)
// tests/test_reference_goldens.py consumes the directory when it exists (oracle on the CPU, CUDA path on the GPU) and skips loudly
// otherwise.  Nothing in this file is used by the product.
package com.fspann.gpu;

import com.fspann.api.ForwardSecureANNSystem;
import com.fspann.common.EncryptedPoint;
import com.fspann.common.QueryResult;
import com.fspann.common.QueryToken;
import com.fspann.common.RocksDBMetadataManager;
import com.fspann.config.SystemConfig;
import com.fspann.crypto.AesGcmCryptoService;
import com.fspann.index.paper.Coding;
import com.fspann.index.paper.GFunctionRegistry;
import com.fspann.index.paper.PartitionedIndexService;
import com.fspann.key.KeyManager;
import com.fspann.key.KeyRotationPolicy;
import com.fspann.key.KeyRotationServiceImpl;
import com.fspann.query.service.QueryServiceImpl;
import io.micrometer.core.instrument.simple.SimpleMeterRegistry;

import java.io.IOException;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.nio.file.Files;
import java.nio.file.Path;
import java.nio.file.Paths;
import java.util.ArrayList;
import java.util.BitSet;
import java.util.LinkedHashMap;
import java.util.List;
import java.util.Map;
import java.util.Random;

public final class GoldenDump {
    static final int N = 3000, DIM = 32, Q = 40, K = 10;
    static final int M = 12, LAMBDA = 2, DIVISIONS = 4, TABLES = 3, REFINE = 64, MAX_GLOBAL = 20000;
    static final long SEED = 13;

    private static final Map<String, String> manifest = new LinkedHashMap<>();
    private static Path out;

    public static void main(String[] args) throws Exception {
        out = Paths.get(args.length > 0 ? args[0] : "tests/golden/reference");
        Files.createDirectories(out);
        Path root = Files.createTempDirectory("fspann-golden");
        Path metaDir = root.resolve("meta"), ptsDir = root.resolve("pts"), ksFile = root.resolve("keys.blob");
        Path seedFile = root.resolve("seed.csv"), cfgFile = root.resolve("cfg.json");
        Files.createDirectories(metaDir);
        Files.createDirectories(ptsDir);
        Files.writeString(seedFile, "");
        Files.writeString(cfgFile, String.format("""
            { "paper": { "enabled": true, "m": %d, "lambda": %d, "divisions": %d, "tables": %d, "seed": %d },
              "runtime": { "refinementLimit": %d, "maxGlobalCandidates": %d, "probeOverride": -1, "hammingPrefilterThreshold": 0 },
              "reencryption": { "enabled": true } }
            """, M, LAMBDA, DIVISIONS, TABLES, SEED, REFINE, MAX_GLOBAL));
        SystemConfig cfg = SystemConfig.load(cfgFile.toString(), true);

        // ---- deterministic SIFT-shaped inputs: 32 cluster centres + noise, rounded to integers 0..255 (what the loaders deliver)
        Random rnd = new Random(7);
        double[][] centres = new double[32][DIM];
        for (double[] c : centres) for (int i = 0; i < DIM; i++) c[i] = 128.0 * rnd.nextDouble();
        List<double[]> base = new ArrayList<>(), queries = new ArrayList<>();
        for (int n = 0; n < N + Q; n++) {
            double[] c = centres[rnd.nextInt(centres.length)], v = new double[DIM];
            for (int i = 0; i < DIM; i++) v[i] = Math.rint(Math.max(0.0, Math.min(255.0, c[i] + 20.0 * rnd.nextGaussian())));
            (n < N ? base : queries).add(v);
        }

        // ---- the stock system, wired like BaseUnifiedIT (the registry initialises itself from the first 1000 inserted vectors, PIS:280-290)
        GFunctionRegistry.reset();
        RocksDBMetadataManager metadata = RocksDBMetadataManager.create(metaDir.toString(), ptsDir.toString());
        KeyManager km = new KeyManager(ksFile.toString());
        KeyRotationServiceImpl keyService = new KeyRotationServiceImpl(km, new KeyRotationPolicy(Integer.MAX_VALUE, Long.MAX_VALUE),
                metaDir.toString(), metadata, null);
        AesGcmCryptoService crypto = new AesGcmCryptoService(new SimpleMeterRegistry(), keyService, metadata);
        keyService.setCryptoService(crypto);
        ForwardSecureANNSystem system = new ForwardSecureANNSystem(cfgFile.toString(), seedFile.toString(), ksFile.toString(), List.of(DIM),
                root, false, metadata, crypto, 64);
        system.setExitOnShutdown(false);
        system.batchInsert(base, DIM);
        system.finalizeForSearch();
        metadata.flush();

        writeDoubles("base", flatten(base), N, DIM);
        writeDoubles("queries", flatten(queries), Q, DIM);

        // ---- GFunctions of the registry (Coding.GFunction alpha / r / omega, Coding:52-97)
        int TD = TABLES * DIVISIONS;
        double[] alpha = new double[TD * M * DIM], r = new double[TD * M], omega = new double[TD * M];
        for (int t = 0; t < TABLES; t++)
            for (int d = 0; d < DIVISIONS; d++) {
                Coding.GFunction g = GFunctionRegistry.get(DIM, t, d);
                int gidx = t * DIVISIONS + d;
                for (int j = 0; j < M; j++) {
                    System.arraycopy(g.alpha[j], 0, alpha, (gidx * M + j) * DIM, DIM);
                    r[gidx * M + j] = g.r[j];
                    omega[gidx * M + j] = g.omega[j];
                }
            }
        writeDoubles("alpha", alpha, TD * M, DIM);
        writeDoubles("r", r, TD, M);
        writeDoubles("omega", omega, TD, M);

        // ---- the encrypted store as persisted (RDB.loadEncryptedPoint, RDB:530-544) and the session keys (KRS:82-88)
        int ctLen = 8 * DIM + 16;
        byte[] iv = new byte[N * 12], ct = new byte[N * ctLen];
        int[] ver = new int[N];
        java.util.TreeSet<Integer> versions = new java.util.TreeSet<>();
        for (int id = 0; id < N; id++) {
            EncryptedPoint ep = metadata.loadEncryptedPoint(Integer.toString(id));
            System.arraycopy(ep.getIv(), 0, iv, id * 12, 12);
            System.arraycopy(ep.getCiphertext(), 0, ct, id * ctLen, ctLen);
            ver[id] = ep.getKeyVersion();
            versions.add(ver[id]);
        }
        writeBytes("store_iv", iv, N, 12);
        writeBytes("store_ct", ct, N, ctLen);
        writeInts("store_key_version", ver, N, 1);
        int[] kv = versions.stream().mapToInt(Integer::intValue).toArray();
        byte[] keys = new byte[kv.length * 32];
        for (int i = 0; i < kv.length; i++) System.arraycopy(keyService.getVersion(kv[i]).getKey().getEncoded(), 0, keys, i * 32, 32);
        writeInts("key_versions", kv, kv.length, 1);
        writeBytes("keys", keys, kv.length, 32);

        // ---- TokenGen -> Route -> Refine for every query
        PartitionedIndexService index = system.getIndexService();
        QueryServiceImpl qs = system.getQueryServiceImpl();
        int W = (M * LAMBDA + 63) / 64;
        long[] codes = new long[Q * TD * W];
        int[] candCount = new int[Q], rawCount = new int[Q], counters = new int[Q * 4], nRet = new int[Q], topIds = new int[Q * K];
        double[] topDist = new double[Q * K];
        List<Integer> candIds = new ArrayList<>(), candScores = new ArrayList<>();
        java.util.Arrays.fill(topIds, -1);
        for (int q = 0; q < Q; q++) {
            QueryToken token = system.createToken(queries.get(q), K, DIM);
            BitSet[][] bc = token.getBitCodes();
            for (int t = 0; t < TABLES; t++)
                for (int d = 0; d < DIVISIONS; d++) {
                    long[] w = bc[t][d].toLongArray();
                    for (int i = 0; i < w.length && i < W; i++) codes[(q * TD + t * DIVISIONS + d) * W + i] = w[i];
                }
            List<PartitionedIndexService.CandidateWithScore> cands = index.lookupCandidatesWithScores(token);
            candCount[q] = cands.size();
            rawCount[q] = index.getLastRawCandidateCount();
            for (PartitionedIndexService.CandidateWithScore c : cands) { candIds.add(Integer.parseInt(c.id())); candScores.add((int) c.hammingDist()); }
            List<QueryResult> res = qs.search(token);
            nRet[q] = res.size();
            for (int i = 0; i < res.size() && i < K; i++) { topIds[q * K + i] = Integer.parseInt(res.get(i).getId()); topDist[q * K + i] = res.get(i).getDistance(); }
            counters[q * 4] = qs.getLastCandTotal(); counters[q * 4 + 1] = qs.getLastCandKept();
            counters[q * 4 + 2] = qs.getLastCandDecrypted(); counters[q * 4 + 3] = qs.getLastReturned();
        }
        writeLongs("codes", codes, Q, TD * W);
        writeInts("cand_count", candCount, Q, 1);
        writeInts("cand_raw_count", rawCount, Q, 1);
        writeInts("cand_ids", candIds.stream().mapToInt(Integer::intValue).toArray(), candIds.size(), 1);
        writeInts("cand_scores", candScores.stream().mapToInt(Integer::intValue).toArray(), candScores.size(), 1);
        writeInts("top_ids", topIds, Q, K);
        writeDoubles("top_dist", topDist, Q, K);
        writeInts("n_ret", nRet, Q, 1);
        writeInts("counters", counters, Q, 4);

        StringBuilder sb = new StringBuilder("{\n  \"params\": {");
        sb.append(String.format("\"N\": %d, \"dim\": %d, \"Q\": %d, \"k\": %d, \"m\": %d, \"lambda\": %d, \"divisions\": %d, \"tables\": %d, \"seed\": %d, "
                + "\"refinementLimit\": %d, \"maxGlobalCandidates\": %d, \"probes\": %d, \"java\": \"%s\"},\n  \"arrays\": {\n",
                N, DIM, Q, K, M, LAMBDA, DIVISIONS, TABLES, SEED, REFINE, MAX_GLOBAL, index.getDefaultMaxProbes(), System.getProperty("java.version")));
        int i = 0;
        for (Map.Entry<String, String> e : manifest.entrySet())
            sb.append("    \"").append(e.getKey()).append("\": ").append(e.getValue()).append(++i < manifest.size() ? ",\n" : "\n");
        sb.append("  }\n}\n");
        Files.writeString(out.resolve("manifest.json"), sb.toString());
        system.shutdown();
        metadata.close();
        System.out.println("golden vectors written to " + out.toAbsolutePath());
    }

    private static double[] flatten(List<double[]> rows) {
        double[] f = new double[rows.size() * rows.get(0).length];
        for (int i = 0; i < rows.size(); i++) System.arraycopy(rows.get(i), 0, f, i * rows.get(0).length, rows.get(0).length);
        return f;
    }

    private static void put(String name, String dtype, ByteBuffer b, int rows, int cols) throws IOException {
        Files.write(out.resolve(name + ".bin"), b.array());
        manifest.put(name, String.format("{\"dtype\": \"%s\", \"shape\": [%d, %d]}", dtype, rows, cols));
    }

    private static void writeDoubles(String name, double[] a, int rows, int cols) throws IOException {
        ByteBuffer b = ByteBuffer.allocate(8 * a.length).order(ByteOrder.LITTLE_ENDIAN);
        for (double v : a) b.putDouble(v);
        put(name, "<f8", b, rows, cols);
    }

    private static void writeLongs(String name, long[] a, int rows, int cols) throws IOException {
        ByteBuffer b = ByteBuffer.allocate(8 * a.length).order(ByteOrder.LITTLE_ENDIAN);
        for (long v : a) b.putLong(v);
        put(name, "<u8", b, rows, cols);
    }

    private static void writeInts(String name, int[] a, int rows, int cols) throws IOException {
        ByteBuffer b = ByteBuffer.allocate(4 * a.length).order(ByteOrder.LITTLE_ENDIAN);
        for (int v : a) b.putInt(v);
        put(name, "<i4", b, rows, cols);
    }

    private static void writeBytes(String name, byte[] a, int rows, int cols) throws IOException {
        put(name, "|u1", ByteBuffer.wrap(a), rows, cols);
    }
}
