/*
 * MhapB200 -- the native-method holder bound by integration/mhapb_jni.c to libmhap_b200.so (include/mhap_b200.h).
 * Lives in edu.umd.marbl.mhap.impl next to GpuMinHashSearch.  The handle is an mhapb_multi*: one JVM, one or several GPUs.
 * NOT compiled in this image (no JDK); tests/test_jni_shim.py checks that every native declared here has its
 * Java_edu_umd_marbl_mhap_impl_MhapB200_* function in the compiled shim.
 */
package edu.umd.marbl.mhap.impl;

import java.nio.ByteBuffer;

final class MhapB200
{
	static
	{
		System.loadLibrary("mhapb_jni");
	}

	private MhapB200()
	{
	}

	/** mhapb_multi_create: one context per listed CUDA device (NCCL communicator inside the library when several). */
	static native long create(int[] devices);

	static native void destroy(long h);

	/** mhapb_host_alloc: a pinned direct buffer to stage reads in (no extra host copy on the way to the GPU). */
	static native ByteBuffer hostAlloc(long bytes);

	static native void hostFree(ByteBuffer buf);

	/** mhapb_multi_store_reset: the constructor arguments of SequenceSketch (impl/SequenceSketch.java:106-116). */
	static native void storeReset(long h, int kmerSize, int numHashes, int orderedKmerSize, int orderedSketchSize, boolean unweighted,
			int minOlapLength);

	/** mhapb_multi_store_add_reads: sketch a batch of reads (both strands) on the GPUs and store them; returns sketches added. */
	static native long storeAddReads(long h, ByteBuffer bases, long[] offsets, long[] ids, int n, boolean bothStrands);

	/** mhapb_multi_store_add_sketches over framed .dat records (byte isFwd, int size, SequenceSketch.getAsByteArray()). */
	static native long storeAddDat(long h, byte[] framedRecords, long idOffset);

	static native long storeSize(long h);

	/** ids / strand flags of all stored sketches (arrays of storeSize entries). */
	static native void storeIds(long h, long[] ids, byte[] isFwd);

	/** One stored sketch as a framed .dat record (index in the numbering of storeIds). */
	static native byte[] storeGetDat(long h, long index);

	/** mhapb_multi_search_self; returns packed mhapb_hit structs (80 bytes each, little-endian), stats[5] = counters. */
	static native byte[] searchSelf(long h, int numMinMatches, int minStoreLength, double maxShift, double acceptScore, long[] stats);

	/** mhapb_search_self restricted to stored sketches [first, first+count): single device only. */
	static native byte[] searchSelfRange(long h, int numMinMatches, int minStoreLength, double maxShift, double acceptScore, long first,
			long count, long[] stats);

	/** mhapb_multi_search_query_reads: query reads sketched forward-only on the GPUs. */
	static native byte[] searchQueryReads(long h, int numMinMatches, int minStoreLength, double maxShift, double acceptScore,
			ByteBuffer bases, long[] offsets, long[] ids, int n, long[] stats);

	/** mhapb_multi_search_query_sketches over framed .dat records. */
	static native byte[] searchQueryDat(long h, int numMinMatches, int minStoreLength, double maxShift, double acceptScore,
			byte[] framedRecords, long idOffset, long[] stats);

	/** mhapb_sketch_to_dat: framed .dat records of a batch of reads (the -p mode, and sketches Java wants as objects). */
	static native byte[] sketchToDat(long h, int kmerSize, int numHashes, int orderedKmerSize, int orderedSketchSize, boolean unweighted,
			int minOlapLength, ByteBuffer bases, long[] offsets, long[] ids, int n, boolean bothStrands);

	/** mhapb_filter_load_text on every device: the text of the -f file; returns the number of repeat k-mers kept. */
	static native long filterLoadText(long h, byte[] text, double filterCutoff, double repeatWeight, double idfScale, int supressNoise,
			boolean noTf, boolean canonical);

	static native void filterClear(long h);
}
