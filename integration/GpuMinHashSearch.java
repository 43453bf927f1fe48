// Java glue stub -- see INTEGRATION.md.  NOT compiled in this image (no JDK).
package edu.umd.marbl.mhap.impl;

final class MhapB200 {
    static { System.loadLibrary("mhapb_jni"); }
    static native long create(int device);
    static native void destroy(long h);
    static native void storeReset(long h, int k, int H, int ok, int os, boolean unweighted, int minOlap);
    static native long storeAddReads(long h, java.nio.ByteBuffer bases, long[] offsets, long[] ids, int n, boolean both);
    static native byte[] searchSelf(long h, int m, int minStore, double maxShift, double accept, long[] stats);
    static native byte[] searchQueryReads(long h, int m, int minStore, double maxShift, double accept,
                                          java.nio.ByteBuffer bases, long[] offsets, long[] ids, int n, long[] stats);
    static native byte[] sketchToDat(long h, java.nio.ByteBuffer bases, long[] offsets, long[] ids, int n, boolean both);
    static native long storeSize(long h);
    /** FrequencyCounts -> device filter: the text of the -f file is parsed by the library (mhapb_filter_load_text). */
    static native long filterLoadText(long h, byte[] text, double filterCutoff, double repeatWeight, double idfScale,
                                      int supressNoise, boolean noTf, boolean canonical);
    static native void filterClear(long h);
}

/** Drop-in for MinHashSearch: same constructor arguments, same getters MhapMain.outputFinalStat reads. */
public final class GpuMinHashSearch extends AbstractMatchSearch {
    private final long ctx;
    private final long[] stats = new long[5];
    // ... k, H, ok, os, m, minStoreLength, maxShift, acceptScore kept from the constructor

    public GpuMinHashSearch(FastaData store, /* same args as MinHashSearch */ ...) {
        super(numThreads, storeResults);
        ctx = MhapB200.create(0);
        MhapB200.storeReset(ctx, k, H, ok, os, repeatWeight < 0.0, minOlapLength);
        // batches of reads straight from FastaData (upper-casing and rc happen on the GPU):
        //   MhapB200.storeAddReads(ctx, bases, offsets, ids, n, true);
    }

    @Override public ArrayList<MatchResult> findMatches() {
        byte[] raw = MhapB200.searchSelf(ctx, numMinMatches, minStoreLength, maxShift, acceptScore, stats);
        ArrayList<MatchResult> out = decode(raw);   // 80-byte records -> new MatchResult(fromId, toId,
        outputResults(out);                         //   new OverlapInfo(score, validCount, a1, a2, b1, b2), fromLen, toLen)
        return out;
    }
    @Override protected boolean addSequence(SequenceSketch s) { throw new UnsupportedOperationException("batched on the GPU"); }
    @Override protected List<MatchResult> findMatches(SequenceSketch q, boolean toSelf) { throw new UnsupportedOperationException(); }
    @Override public int size() { return (int) MhapB200.storeSize(ctx); }
    public long getNumberElementsProcessed() { return stats[0]; }
    public long getNumberSequencesHit() { return stats[1]; }
    public long getNumberSequencesFullyCompared() { return stats[2]; }
    // getStoredForwardSequenceIds / getStoredSequenceHash: via sketchToDat + SequenceSketch.fromByteStream
}
