// Source only: this repository's image has no JDK, so these files are not compiled here.  They are the reference-side binding a
// maintainer of Mehran-Memon/fspann-query-system would add (see INTEGRATION.md); the same C ABI is exercised from C++
// (include/fspann_host.hpp) and Python (fspann_query_system_b200/gpu.py) by the test-suite.
package com.fspann.gpu;

import java.lang.foreign.*;
import java.lang.invoke.MethodHandle;
import static java.lang.foreign.ValueLayout.*;

/** Thin downcall layer over include/fspann_gpu.h.  One instance per GPU. */
public final class FspannGpu implements AutoCloseable {
    private static final Linker L = Linker.nativeLinker();
    private static final SymbolLookup LIB = SymbolLookup.libraryLookup("libfspann_gpu.so", Arena.global());
    private static MethodHandle h(String n, FunctionDescriptor d) { return L.downcallHandle(LIB.find(n).orElseThrow(), d); }

    private static final MethodHandle CTX_CREATE   = h("fspann_ctx_create",   FunctionDescriptor.of(JAVA_INT, JAVA_INT, ADDRESS));
    private static final MethodHandle CTX_DESTROY  = h("fspann_ctx_destroy",  FunctionDescriptor.ofVoid(ADDRESS));
    private static final MethodHandle LAST_ERROR   = h("fspann_last_error",   FunctionDescriptor.of(ADDRESS, ADDRESS));
    private static final MethodHandle ROUTING_UP   = h("fspann_routing_upload", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT,
            JAVA_INT, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle STORE_UP     = h("fspann_store_upload", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle STORE_UPDATE = h("fspann_store_update", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle KEYS_SET     = h("fspann_keys_set",     FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
    private static final MethodHandle KEYS_RETIRE  = h("fspann_keys_retire",  FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
    private static final MethodHandle SEARCH       = h("fspann_search_batch", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, JAVA_INT, JAVA_INT,
            JAVA_LONG, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle SEARCH_TOK   = h("fspann_search_tokens", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT,
            JAVA_LONG, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle COMM_ID      = h("fspann_comm_unique_id", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    private static final MethodHandle COMM_INIT    = h("fspann_comm_init", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS));
    private static final MethodHandle STORE_SHARD  = h("fspann_store_upload_shard", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_LONG, JAVA_INT,
            ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle SHARDED      = h("fspann_sharded_search_batch", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, JAVA_INT, JAVA_INT,
            JAVA_LONG, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle TOUCHED      = h("fspann_touched_fetch", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_INT));

    private final MemorySegment ctx;

    public FspannGpu(int device) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment out = a.allocate(ADDRESS);
            int rc = (int) CTX_CREATE.invokeExact(device, out);
            if (rc != 0) throw new IllegalStateException("no CUDA device: the FSPANN hot path has no CPU fallback (rc=" + rc + ")");
            ctx = out.get(ADDRESS, 0);
        }
    }

    /** Maps the ABI's error classes back onto the exceptions the reference throws on this path. */
    private void check(int rc) throws Throwable {
        if (rc == 0) return;
        String msg = ((MemorySegment) LAST_ERROR.invokeExact(ctx)).reinterpret(512).getUtf8String(0);
        switch (rc) {
            case -1 -> throw new IllegalArgumentException(msg);   // FSPANN_E_ARG
            case -2 -> throw new IllegalStateException(msg);      // FSPANN_E_STATE ("Index not finalized", PIS:594)
            default -> throw new RuntimeException("CUDA failure " + rc + ": " + msg);
        }
    }

    public void keysSet(int version, byte[] key32) throws Throwable {
        try (Arena a = Arena.ofConfined()) { check((int) KEYS_SET.invokeExact(ctx, version, a.allocateArray(JAVA_BYTE, key32))); }
    }
    public void keysRetire(int version) throws Throwable { check((int) KEYS_RETIRE.invokeExact(ctx, version)); }

    /** queries: Q*dim doubles; returns ids/dists/nRet/counters in caller-provided segments (all host memory, copied by the library). */
    public void searchBatch(long q, MemorySegment queries, int k, int probes, long hardCap, int refinementLimit, int hammingThreshold,
                            MemorySegment idsOut, MemorySegment distOut, MemorySegment nRetOut, MemorySegment countersOut) throws Throwable {
        check((int) SEARCH.invokeExact(ctx, q, queries, k, probes, hardCap, refinementLimit, hammingThreshold, idsOut, distOut, nRetOut, countersOut));
    }
    /** QueryServiceImpl.search on the tokens' own codes (PIS:600): codes = q*T*D*W longs (BitSet.toLongArray words), queries = decrypted payloads. */
    public void searchTokens(long q, MemorySegment codes, MemorySegment queries, int k, int probes, long hardCap, int refinementLimit, int hammingThreshold,
                             MemorySegment idsOut, MemorySegment distOut, MemorySegment nRetOut, MemorySegment countersOut) throws Throwable {
        check((int) SEARCH_TOK.invokeExact(ctx, q, codes, queries, k, probes, hardCap, refinementLimit, hammingThreshold, idsOut, distOut, nRetOut, countersOut));
    }

    // ---- database-sharded deployment (BASELINE config 4): one FspannGpu per GPU, the NCCL collectives run inside the library
    /** Any one participant draws the 128-byte communicator id and hands it to the others. */
    public static byte[] commUniqueId() throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment id = a.allocate(128);
            if ((int) COMM_ID.invokeExact(id) != 0) throw new IllegalStateException("NCCL unavailable (libnccl.so.2 not loadable; set FSPANN_NCCL_LIB)");
            return id.toArray(JAVA_BYTE);
        }
    }
    /** Collective: every participant calls it with the same id and its own rank; rank r holds the r-th contiguous id range of the store. */
    public void commInit(int nRanks, int rank, byte[] id) throws Throwable {
        try (Arena a = Arena.ofConfined()) { check((int) COMM_INIT.invokeExact(ctx, nRanks, rank, a.allocateArray(JAVA_BYTE, id))); }
    }
    public void storeUploadShard(long idBase, long n, long nGlobal, int dim, MemorySegment iv, MemorySegment ct, MemorySegment keyVersion) throws Throwable {
        check((int) STORE_SHARD.invokeExact(ctx, idBase, n, nGlobal, dim, iv, ct, keyVersion));
    }
    /** Collective: the same batch on every participant, the same (unsharded-identical) result on every participant. */
    public void shardedSearchBatch(long q, MemorySegment queries, int k, int probes, long hardCap, int refinementLimit,
                                   MemorySegment idsOut, MemorySegment distOut, MemorySegment nRetOut, MemorySegment countersOut) throws Throwable {
        check((int) SHARDED.invokeExact(ctx, q, queries, k, probes, hardCap, refinementLimit, idsOut, distOut, nRetOut, countersOut));
    }
    // routingUpload / storeUpload / storeUpdate / touchedFetch follow the same pattern.

    @Override public void close() { try { CTX_DESTROY.invokeExact(ctx); } catch (Throwable ignored) { } }

    private static final MethodHandle MIGRATE = h("fspann_migrate", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS, JAVA_INT,
            ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    /** KRS:215-289 with the AES-GCM work on the device.  ids / freshIvs / outputs are off-heap segments of n, n*12, n, n*12, n*(8*dim+16) bytes. */
    public long migrate(long n, MemorySegment ids, MemorySegment freshIvs, int targetVersion, MemorySegment done, MemorySegment ivOut,
                        MemorySegment ctOut) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment cnt = a.allocate(JAVA_LONG);
            check((int) MIGRATE.invokeExact(ctx, n, ids, freshIvs, targetVersion, done, ivOut, ctOut, cnt));
            return cnt.get(JAVA_LONG, 0);                       // ReencryptReport.reencrypted
        }
    }

}
