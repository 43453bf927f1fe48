// PinOracle.java -- run on a box with a JDK and the reference's jar to PIN the oracle (DESIGN.md 2: parity is unpinned
// because no JVM exists in the build image).  Not compiled or run here.
//
//   javac -cp mhap-2.1.3.jar PinOracle.java
//   java  -cp mhap-2.1.3.jar:. PinOracle tests/golden/pin_cases.txt > jvm_vectors.txt
//   python tests/golden/check_against_jvm.py jvm_vectors.txt        # diffs against the oracle (oracle/mhap_oracle.c)
//
// Input: one case per line (written by tests/golden/check_against_jvm.py --write-cases):
//   H <k> <seq>                          -> computeSequenceHashesLong / computeSequenceHashes      (HashUtils.java:213-258)
//   M <k> <H> <repeatWeight> <seq>       -> MinHashSketch without a filter                         (MinHashSketch.java:51-179)
//   B <ok> <S> <seq>                     -> BottomOverlapSketch.getAsByteArray                      (BottomOverlapSketch.java:525-585)
//   O <ok> <S> <maxShift> <seqA> <seqB>  -> getOverlapInfo                                          (BottomOverlapSketch.java:592-630)
//   F <k> <H> <repeatWeight> <cutoff> <supressNoise> <noTf> <range> <filterFile> <seq>  -> MinHashSketch with FrequencyCounts
// Output: the same tag followed by the reference's numbers, one line per case, in the order of the input.
import java.io.BufferedReader;
import java.io.FileReader;
import java.nio.ByteBuffer;

import edu.umd.marbl.mhap.impl.OverlapInfo;
import edu.umd.marbl.mhap.sketch.BottomOverlapSketch;
import edu.umd.marbl.mhap.sketch.FrequencyCounts;
import edu.umd.marbl.mhap.sketch.HashUtils;
import edu.umd.marbl.mhap.sketch.MinHashSketch;

public final class PinOracle {
    static String join(long[] a) { StringBuilder s = new StringBuilder(); for (long v : a) s.append(' ').append(v); return s.toString(); }
    static String join(int[] a) { StringBuilder s = new StringBuilder(); for (int v : a) s.append(' ').append(v); return s.toString(); }

    public static void main(String[] args) throws Exception {
        try (BufferedReader in = new BufferedReader(new FileReader(args[0]))) {
            for (String line = in.readLine(); line != null; line = in.readLine()) {
                if (line.isEmpty() || line.charAt(0) == '#') continue;
                String[] f = line.split(" ");
                try {
                    switch (f[0]) {
                    case "H": {
                        int k = Integer.parseInt(f[1]);
                        System.out.println("H64" + join(HashUtils.computeSequenceHashesLong(f[2], k, 0, false)));
                        System.out.println("H64C" + join(HashUtils.computeSequenceHashesLong(f[2], k, 0, true)));
                        System.out.println("H32" + join(HashUtils.computeSequenceHashes(f[2], k, false)));
                        break;
                    }
                    case "M": {
                        MinHashSketch m = new MinHashSketch(f[4], Integer.parseInt(f[1]), Integer.parseInt(f[2]), null, false, Double.parseDouble(f[3]));
                        System.out.println("M" + join(m.getMinHashArray()));
                        break;
                    }
                    case "B": {
                        // getAsByteArray (:561-585): int seqLength, int kmerSize, int n, then n x (int hash, int pos), big-endian
                        ByteBuffer b = ByteBuffer.wrap(new BottomOverlapSketch(f[3], Integer.parseInt(f[1]), Integer.parseInt(f[2]), false).getAsByteArray());
                        StringBuilder s = new StringBuilder("B");
                        while (b.remaining() >= 4) s.append(' ').append(b.getInt());
                        System.out.println(s);
                        break;
                    }
                    case "O": {
                        int ok = Integer.parseInt(f[1]), S = Integer.parseInt(f[2]);
                        BottomOverlapSketch a = new BottomOverlapSketch(f[4], ok, S, false), c = new BottomOverlapSketch(f[5], ok, S, false);
                        OverlapInfo o = a.getOverlapInfo(c, Double.parseDouble(f[3]));
                        System.out.println("O " + o.a1 + " " + o.a2 + " " + o.b1 + " " + o.b2 + " " + (long) o.rawScore + " " + Double.doubleToLongBits(o.score));
                        break;
                    }
                    case "F": {
                        double rw = Double.parseDouble(f[3]);
                        double offset = (rw >= 0.0 && rw < 1.0) ? rw : 0.0;          // main/MhapMain.java:348-350
                        FrequencyCounts fc;
                        try (BufferedReader bf = new BufferedReader(new FileReader(f[8]))) {
                            fc = new FrequencyCounts(bf, Double.parseDouble(f[4]), offset, Integer.parseInt(f[5]), f[6].equals("1"), 1,
                                                     Double.parseDouble(f[7]), true);
                        }
                        MinHashSketch m = new MinHashSketch(f[9], Integer.parseInt(f[1]), Integer.parseInt(f[2]), fc, false, rw);
                        System.out.println("F" + join(m.getMinHashArray()));
                        break;
                    }
                    default:
                        System.out.println("? " + f[0]);
                    }
                } catch (edu.umd.marbl.mhap.sketch.ZeroNGramsFoundException e) {
                    System.out.println(f[0] + " ZERO");
                }
            }
        }
    }
}
