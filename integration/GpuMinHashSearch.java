/*
 * GpuMinHashSearch -- drop-in for impl/MinHashSearch.java behind AbstractMatchSearch's seams, backed by libmhap_b200.so.
 *
 * Same role as MinHashSearch (impl/MinHashSearch.java:45-306): the constructor sketches and indexes the store file,
 * findMatches() is the self overlap, findMatches(SequenceSketchStreamer) the store-vs-query mode, and the getters are the
 * ones MhapMain.outputFinalStat prints (main/MhapMain.java:572-590).  The per-sequence thread pools of
 * AbstractMatchSearch (:67-117,121-199,203-285) are replaced by batched calls; the per-sketch abstract methods are still
 * implemented (through the .dat record form of a sketch) so that nothing of the base class is left dangling.
 *
 * Must live in package edu.umd.marbl.mhap.impl: MatchResult's constructor is protected (impl/MatchResult.java:46).
 * NOT compiled in this image (no JDK, no Maven, Guava/fastutil un-vendored); see INTEGRATION.md and MhapMain.patch.
 */
package edu.umd.marbl.mhap.impl;

import java.io.ByteArrayInputStream;
import java.io.ByteArrayOutputStream;
import java.io.DataInputStream;
import java.io.DataOutputStream;
import java.io.IOException;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.nio.charset.StandardCharsets;
import java.nio.file.Files;
import java.nio.file.Paths;
import java.util.ArrayList;
import java.util.List;
import java.util.concurrent.atomic.AtomicLong;

import edu.umd.marbl.mhap.utils.ReadBuffer;

public final class GpuMinHashSearch extends AbstractMatchSearch
{
	/** bytes of one packed mhapb_hit (include/mhap_b200.h) */
	private static final int HIT_BYTES = 80;
	/** reads per batch handed to the GPUs: ~256 MB of bases at 10 kbp */
	private static final int BATCH_READS = 25000;
	private static final long BATCH_BASES = 256L << 20;

	private final long handle;
	private final int kmerSize, numHashes, orderedKmerSize, orderedSketchSize, minOlapLength;
	private final boolean unweighted;
	private final int numMinMatches, minStoreLength;
	private final double maxShift, acceptScore;

	private final AtomicLong numberElementsProcessed = new AtomicLong();
	private final AtomicLong numberSequencesHit = new AtomicLong();
	private final AtomicLong numberSequencesFullyCompared = new AtomicLong();
	private final AtomicLong matchesProcessedGpu = new AtomicLong();
	private final AtomicLong sequencesSearchedGpu = new AtomicLong();
	private final AtomicLong searchTime = new AtomicLong();
	private int fastaProcessed = 0;

	/**
	 * @param storeFile  the -s file: FASTA (plain, the reference's FastaData reads it) or .dat
	 * @param devices    CUDA ordinals to shard the reads over, e.g. {0} or {0,1,2,3,4,5,6,7}
	 * @param filterText bytes of the -f k-mer filter file, or null (main/MhapMain.java:340-372)
	 */
	public GpuMinHashSearch(String storeFile, int[] devices, int minOlapLength, int kmerSize, int numHashes, int orderedKmerSize,
			int orderedSketchSize, double repeatWeight, byte[] filterText, double filterCutoff, double idfScale, int supressNoise,
			boolean noTf, boolean doReverseCompliment, int numMinMatches, int numThreads, boolean storeResults, int minStoreLength,
			double maxShift, double acceptScore) throws IOException
	{
		super(numThreads, storeResults);

		this.kmerSize = kmerSize;
		this.numHashes = numHashes;
		this.orderedKmerSize = orderedKmerSize;
		this.orderedSketchSize = orderedSketchSize;
		this.minOlapLength = minOlapLength;
		this.unweighted = repeatWeight < 0.0;
		this.numMinMatches = numMinMatches;
		this.minStoreLength = minStoreLength;
		this.maxShift = maxShift;
		this.acceptScore = acceptScore;

		this.handle = MhapB200.create(devices);
		MhapB200.storeReset(this.handle, kmerSize, numHashes, orderedKmerSize, orderedSketchSize, this.unweighted, minOlapLength);
		if (filterText != null)
			MhapB200.filterLoadText(this.handle, filterText, filterCutoff, repeatWeight, idfScale, supressNoise, noTf, doReverseCompliment);

		if (storeFile.endsWith(".dat"))
		{
			// SequenceSketchStreamer.readFromBinary (impl/SequenceSketchStreamer.java:278-320): the file IS the framed records
			byte[] records = Files.readAllBytes(Paths.get(storeFile));
			MhapB200.storeAddDat(this.handle, records, 0);
		}
		else
		{
			// data.enqueueFullFile + addData (impl/MinHashSearch.java:80,95): batches of reads, both strands sketched on the GPU
			FastaData fasta = new FastaData(storeFile, 0);
			ReadBatch batch = new ReadBatch();
			while (batch.fill(fasta))
				MhapB200.storeAddReads(this.handle, batch.bases, batch.offsets(), batch.ids(), batch.size(), true);
			batch.free();
			this.fastaProcessed = fasta.getNumberProcessed();
		}

		System.err.println("Stored " + size() + " sequences in the index.");
	}

	/** Number of FASTA records read for the store (MhapMain uses it as the id offset of the query files, :462). */
	public int getFastaProcessed()
	{
		return this.fastaProcessed;
	}

	public void close()
	{
		MhapB200.destroy(this.handle);
	}

	// ---- the three search entry points --------------------------------------------------------------------------

	/** findMatches() to self (impl/AbstractMatchSearch.java:121-199): one call, all forward sketches query the index. */
	@Override
	public ArrayList<MatchResult> findMatches()
	{
		long startTime = System.nanoTime();
		long[] stats = new long[5];
		byte[] raw = MhapB200.searchSelf(this.handle, this.numMinMatches, this.minStoreLength, this.maxShift, this.acceptScore, stats);
		this.searchTime.getAndAdd(System.nanoTime() - startTime);
		return finish(decode(raw, null), stats);
	}

	/**
	 * findMatches(SequenceSketchStreamer) (impl/AbstractMatchSearch.java:203-285).  The streamer hands out SequenceSketch
	 * objects (it owns the FASTA reader), so this generic form takes them as they come -- forward only, :225 -- and searches
	 * them in batches through their .dat record form.  findMatches(String, long) below is the fast path for FASTA queries:
	 * it lets the GPUs do the sketching too.
	 */
	@Override
	public ArrayList<MatchResult> findMatches(final SequenceSketchStreamer data) throws IOException
	{
		ArrayList<MatchResult> all = new ArrayList<MatchResult>();
		ReadBuffer buf = new ReadBuffer();
		ByteArrayOutputStream records = new ByteArrayOutputStream();
		DataOutputStream dos = new DataOutputStream(records);
		ArrayList<String> headers = new ArrayList<String>();
		int inBatch = 0;

		SequenceSketch sketch = data.dequeue(true, buf);
		while (sketch != null || inBatch > 0)
		{
			if (sketch != null)
			{
				frame(dos, sketch);
				inBatch++;
				sketch = data.dequeue(true, buf);
			}
			if (inBatch > 0 && (sketch == null || inBatch >= BATCH_READS))
			{
				dos.flush();
				long startTime = System.nanoTime();
				long[] stats = new long[5];
				byte[] raw = MhapB200.searchQueryDat(this.handle, this.numMinMatches, this.minStoreLength, this.maxShift, this.acceptScore,
						records.toByteArray(), 0, stats);
				this.searchTime.getAndAdd(System.nanoTime() - startTime);
				all.addAll(finish(decode(raw, null), stats));
				records.reset();
				inBatch = 0;
			}
		}
		flushOutput();
		return all;
	}

	/** Store-vs-query for a FASTA query file: reads are sketched forward-only on the GPUs (mhapb_multi_search_query_reads). */
	public ArrayList<MatchResult> findMatches(String queryFastaFile, long idOffset) throws IOException
	{
		ArrayList<MatchResult> all = new ArrayList<MatchResult>();
		FastaData fasta = new FastaData(queryFastaFile, idOffset);
		ReadBatch batch = new ReadBatch();
		while (batch.fill(fasta))
		{
			long startTime = System.nanoTime();
			long[] stats = new long[5];
			byte[] raw = MhapB200.searchQueryReads(this.handle, this.numMinMatches, this.minStoreLength, this.maxShift, this.acceptScore,
					batch.bases, batch.offsets(), batch.ids(), batch.size(), stats);
			this.searchTime.getAndAdd(System.nanoTime() - startTime);
			all.addAll(finish(decode(raw, batch), stats));
		}
		batch.free();
		flushOutput();
		return all;
	}

	// ---- the per-sketch seams of AbstractMatchSearch ------------------------------------------------------------

	/** addSequence (impl/MinHashSearch.java:101-147) for a sketch computed elsewhere: stored through its .dat record. */
	@Override
	protected boolean addSequence(SequenceSketch currHash)
	{
		try
		{
			ByteArrayOutputStream records = new ByteArrayOutputStream();
			DataOutputStream dos = new DataOutputStream(records);
			frame(dos, currHash);
			dos.flush();
			return MhapB200.storeAddDat(this.handle, records.toByteArray(), 0) == 1;
		}
		catch (IOException e)
		{
			throw new MhapRuntimeException(e);
		}
	}

	/** findMatches(sketch, toSelf) (impl/MinHashSearch.java:150-251) for one sketch; the batched forms above are the fast ones. */
	@Override
	protected List<MatchResult> findMatches(SequenceSketch seqHashes, boolean toSelf)
	{
		long[] stats = new long[5];
		byte[] raw;
		if (toSelf)
		{
			// a stored sequence queries the index under the self-search id rules (:200,215-225): single device only
			long index = storedIndexOf(seqHashes.getSequenceId());
			raw = MhapB200.searchSelfRange(this.handle, this.numMinMatches, this.minStoreLength, this.maxShift, this.acceptScore, index, 1, stats);
		}
		else
		{
			try
			{
				ByteArrayOutputStream records = new ByteArrayOutputStream();
				DataOutputStream dos = new DataOutputStream(records);
				frame(dos, seqHashes);
				dos.flush();
				raw = MhapB200.searchQueryDat(this.handle, this.numMinMatches, this.minStoreLength, this.maxShift, this.acceptScore,
						records.toByteArray(), 0, stats);
			}
			catch (IOException e)
			{
				throw new MhapRuntimeException(e);
			}
		}
		ArrayList<MatchResult> matches = decode(raw, null);
		accumulate(stats);
		return matches;
	}

	@Override
	public List<SequenceId> getStoredForwardSequenceIds()
	{
		int n = size();
		long[] ids = new long[n];
		byte[] fwd = new byte[n];
		MhapB200.storeIds(this.handle, ids, fwd);
		ArrayList<SequenceId> seqIds = new ArrayList<SequenceId>(n / 2 + 1);
		for (int i = 0; i < n; i++)
			if (fwd[i] != 0)
				seqIds.add(new SequenceId(ids[i], true));
		return seqIds;
	}

	@Override
	public SequenceSketch getStoredSequenceHash(SequenceId id)
	{
		long index = storedIndexOf(id);
		byte[] record = MhapB200.storeGetDat(this.handle, index);
		try
		{
			// framing: byte isFwd, int payload size (impl/SequenceSketchStreamer.java:291-303), then SequenceSketch.fromByteStream
			DataInputStream in = new DataInputStream(new ByteArrayInputStream(record));
			in.readByte();
			in.readInt();
			return SequenceSketch.fromByteStream(in, 0);
		}
		catch (IOException e)
		{
			throw new MhapRuntimeException(e);
		}
	}

	@Override
	public int size()
	{
		return (int) MhapB200.storeSize(this.handle);
	}

	// ---- the getters MhapMain.outputFinalStat reads (main/MhapMain.java:572-590) --------------------------------

	@Override
	public long getMatchesProcessed()
	{
		return this.matchesProcessedGpu.get();
	}

	@Override
	public long getNumberSequencesSearched()
	{
		return this.sequencesSearchedGpu.get();
	}

	public double getMinHashSearchTime()
	{
		return this.searchTime.longValue() * 1.0e-9;
	}

	public long getNumberElementsProcessed()
	{
		return this.numberElementsProcessed.get();
	}

	public long getNumberSequencesFullyCompared()
	{
		return this.numberSequencesFullyCompared.get();
	}

	public long getNumberSequencesHit()
	{
		return this.numberSequencesHit.get();
	}

	// ---- helpers ------------------------------------------------------------------------------------------------

	private void accumulate(long[] stats)
	{
		this.numberElementsProcessed.getAndAdd(stats[0]);
		this.numberSequencesHit.getAndAdd(stats[1]);
		this.numberSequencesFullyCompared.getAndAdd(stats[2]);
		this.matchesProcessedGpu.getAndAdd(stats[3]);
		this.sequencesSearchedGpu.getAndAdd(stats[4]);
	}

	/** counters + output, as the worker loops of AbstractMatchSearch do every NUM_ELEMENTS_PER_OUTPUT matches (:158-175) */
	private ArrayList<MatchResult> finish(ArrayList<MatchResult> matches, long[] stats)
	{
		accumulate(stats);
		outputResults(matches);   // no-op when storeResults is set (:318)
		flushOutput();
		return matches;
	}

	/** byte isFwd, int size, payload: the record framing of SequenceSketchStreamer.writeToBinary (:349-360) */
	private static void frame(DataOutputStream dos, SequenceSketch sketch) throws IOException
	{
		byte[] payload = sketch.getAsByteArray();
		dos.writeBoolean(sketch.getSequenceId().isForward());
		dos.writeInt(payload.length);
		dos.write(payload);
	}

	private long storedIndexOf(SequenceId id)
	{
		int n = size();
		long[] ids = new long[n];
		byte[] fwd = new byte[n];
		MhapB200.storeIds(this.handle, ids, fwd);
		for (int i = 0; i < n; i++)
			if (ids[i] == id.getHeaderId() && (fwd[i] != 0) == id.isForward())
				return i;
		throw new MhapRuntimeException("Sequence " + id + " is not stored.");
	}

	/**
	 * Packed mhapb_hit records -> MatchResult (impl/MatchResult.java:46-65 applies the strand flip itself).  Layout, little-endian:
	 * long fromId, long toId, int fromFwd, int toFwd, int hitCount, int a1, a2, b1, b2, int validCount, int intersect, int kmin,
	 * int fromLen, int toLen, double score, int accepted, int pad.  names: FASTA headers of a query batch for --store-full-id.
	 */
	private static ArrayList<MatchResult> decode(byte[] raw, ReadBatch names)
	{
		int n = raw == null ? 0 : raw.length / HIT_BYTES;
		ArrayList<MatchResult> out = new ArrayList<MatchResult>(n);
		if (n == 0)
			return out;
		ByteBuffer b = ByteBuffer.wrap(raw).order(ByteOrder.LITTLE_ENDIAN);
		for (int i = 0; i < n; i++)
		{
			int p = i * HIT_BYTES;
			long fromId = b.getLong(p);
			long toId = b.getLong(p + 8);
			boolean fromFwd = b.getInt(p + 16) != 0;
			boolean toFwd = b.getInt(p + 20) != 0;
			int a1 = b.getInt(p + 28), a2 = b.getInt(p + 32), b1 = b.getInt(p + 36), b2 = b.getInt(p + 40);
			int validCount = b.getInt(p + 44);
			int fromLen = b.getInt(p + 56), toLen = b.getInt(p + 60);
			double score = b.getDouble(p + 64);
			boolean accepted = b.getInt(p + 72) != 0;
			if (!accepted)
				continue;
			// rawScore = number of valid shared ordered k-mers (sketch/BottomOverlapSketch.java:613,629)
			OverlapInfo overlap = new OverlapInfo(score, (double) validCount, a1, a2, b1, b2);
			String fromName = names == null ? null : names.headerOf(fromId);
			SequenceId from = fromName == null ? new SequenceId(fromId, fromFwd) : new SequenceId(fromId, fromFwd, fromName);
			out.add(new MatchResult(from, new SequenceId(toId, toFwd), overlap, fromLen, toLen));
		}
		return out;
	}

	/** A batch of reads in the layout the library takes: one pinned direct buffer of bases + offsets + ids. */
	private static final class ReadBatch
	{
		ByteBuffer bases = MhapB200.hostAlloc(BATCH_BASES);
		private final long[] off = new long[BATCH_READS + 1];
		private final long[] id = new long[BATCH_READS];
		private final String[] header = new String[BATCH_READS];
		private int n = 0;
		private Sequence pending = null;

		/** Next batch from the reader; false when the file is exhausted.  FastaData upper-cases (:194); the GPU does it again, harmlessly. */
		boolean fill(FastaData fasta) throws IOException
		{
			this.n = 0;
			this.bases.clear();
			this.off[0] = 0;
			Sequence seq = this.pending != null ? this.pending : fasta.dequeue();
			this.pending = null;
			while (seq != null)
			{
				byte[] chars = seq.getSquenceString().getBytes(StandardCharsets.ISO_8859_1);
				if (this.n > 0 && (this.n >= BATCH_READS || this.bases.position() + chars.length > this.bases.capacity()))
				{
					this.pending = seq;
					break;
				}
				if (chars.length > this.bases.capacity())
					throw new MhapRuntimeException("Sequence longer than the staging buffer.");
				this.bases.put(chars);
				this.id[this.n] = seq.getId().getHeaderId();
				this.header[this.n] = SequenceId.STORE_FULL_ID ? seq.getId().getHeader() : null;
				this.n++;
				this.off[this.n] = this.bases.position();
				seq = fasta.dequeue();
			}
			return this.n > 0;
		}

		int size()
		{
			return this.n;
		}

		long[] offsets()
		{
			return java.util.Arrays.copyOf(this.off, this.n + 1);
		}

		long[] ids()
		{
			return java.util.Arrays.copyOf(this.id, this.n);
		}

		String headerOf(long seqId)
		{
			for (int i = 0; i < this.n; i++)
				if (this.id[i] == seqId)
					return this.header[i];
			return null;
		}

		void free()
		{
			MhapB200.hostFree(this.bases);
			this.bases = null;
		}
	}
}
