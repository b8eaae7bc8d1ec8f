// Source only: this repository's image has no JDK, so these files are not compiled here.  They are the reference-side binding a
// maintainer of Mehran-Memon/fspann-query-system would add (see INTEGRATION.md); the same C ABI is exercised from C++
// (include/fspann_host.hpp) and Python (fspann_query_system_b200/gpu.py) by the test-suite.
package com.fspann.gpu;

import com.fspann.common.*;
import com.fspann.config.SystemConfig;
import com.fspann.crypto.CryptoService;
import com.fspann.index.paper.PartitionedIndexService;
import com.fspann.query.service.QueryService;

import java.lang.foreign.*;
import java.util.*;
import static java.lang.foreign.ValueLayout.*;

public final class GpuQueryServiceImpl implements QueryService {
    private final FspannGpu gpu; private final PartitionedIndexService index; private final CryptoService crypto;
    private final KeyLifeCycleService keys; private final SystemConfig cfg;
    private volatile int lastCandTotal, lastCandKept, lastCandDecrypted, lastReturned;

    @Override public List<QueryResult> search(QueryToken token) {           // QSI:101
        if (token == null) return Collections.emptyList();                   // QSI:102
        return searchBatch(List.of(token)).get(0);
    }

    public List<List<QueryResult>> searchBatch(List<QueryToken> tokens) {
        if (!index.isFrozen()) throw new IllegalStateException("Index not finalized");       // PIS:594
        int q = tokens.size(), dim = tokens.get(0).getDimension(), k = tokens.get(0).getTopK();
        try (Arena a = Arena.ofConfined()) {
            var paper = cfg.getPaper();
            int td = paper.getTables() * paper.getDivisions(), w = (paper.getM() * paper.getLambda() + 63) / 64;
            MemorySegment qs = a.allocateArray(JAVA_DOUBLE, (long) q * dim);
            MemorySegment codes = a.allocateArray(JAVA_LONG, (long) q * td * w);   // zero-filled: BitSet.toLongArray drops trailing zero words
            for (int i = 0; i < q; i++) {                                    // QSI:124-135: decrypt the query inside the trusted component
                QueryToken t = tokens.get(i);
                KeyVersion kv; try { kv = keys.getVersion(t.getVersion()); } catch (Throwable e) { kv = keys.getCurrentVersion(); }
                double[] v = crypto.decryptQuery(t.getEncryptedQuery(), t.getIv(), kv.getKey());
                MemorySegment.copy(v, 0, qs, JAVA_DOUBLE, (long) i * dim * 8, dim);
                BitSet[][] bc = t.getBitCodes();                             // Route runs on the token's OWN codes (PIS:600): no second TokenGen
                if (bc == null) throw new IllegalStateException("MSANNP violation: QueryToken missing BitSet codes");   // PIS:604-606
                for (int tt = 0; tt < paper.getTables(); tt++)
                    for (int d = 0; d < paper.getDivisions(); d++) {
                        long[] words = bc[tt][d].toLongArray();
                        for (int x = 0; x < words.length && x < w; x++) codes.setAtIndex(JAVA_LONG, ((long) i * td + (long) tt * paper.getDivisions() + d) * w + x, words[x]);
                    }
            }
            MemorySegment ids = a.allocateArray(JAVA_INT, (long) q * k), dist = a.allocateArray(JAVA_DOUBLE, (long) q * k),
                          nRet = a.allocateArray(JAVA_INT, q), cnt = a.allocateArray(JAVA_LONG, (long) q * 6);
            var rt = cfg.getRuntime();
            try {
                gpu.searchTokens(q, codes, qs, k, index.effectiveMaxProbes(), Math.max(rt.getMaxGlobalCandidates(), rt.getRefinementLimit()),
                                 getEffectiveRefinementLimit(rt.getRefinementLimit()), rt.getHammingPrefilterThreshold(), ids, dist, nRet, cnt);
            } finally {
                index.clearProbeOverride();                                  // QSI:342-346 (finally)
            }
            List<List<QueryResult>> out = new ArrayList<>(q);
            for (int i = 0; i < q; i++) {
                int n = nRet.getAtIndex(JAVA_INT, i);
                List<QueryResult> r = new ArrayList<>(n);
                for (int j = 0; j < n; j++)
                    r.add(new QueryResult(Integer.toString(ids.getAtIndex(JAVA_INT, (long) i * k + j)), dist.getAtIndex(JAVA_DOUBLE, (long) i * k + j)));
                out.add(r);
            }
            long base = (long) (q - 1) * 6;                                  // getLast* describe the last query, as in the reference
            lastCandTotal = (int) cnt.getAtIndex(JAVA_LONG, base); lastCandKept = (int) cnt.getAtIndex(JAVA_LONG, base + 1);
            lastCandDecrypted = (int) cnt.getAtIndex(JAVA_LONG, base + 2); lastReturned = (int) cnt.getAtIndex(JAVA_LONG, base + 3);
            // reencTracker.record(touched) (QSI:348-350): gpu.touchedFetch(bitmap, clear=true) -> ids
            return out;
        } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); }
    }
}
