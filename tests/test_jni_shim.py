"""CPU suite: the reference-side glue under integration/ is real source, not a sketch.

There is no JDK in this image, so the JNI shim is compiled against tests/cpp/jni_stub/jni.h (the JNI names and signatures
the shim uses, nothing else) and checked three ways: it compiles warning-free, it links against libmhap_b200.so into a
shared object, and every `static native` declared in integration/MhapB200.java has its Java_..._MhapB200_<name> symbol in
that object (and vice versa).  The GPU half (tests/test_gpu_jni_shim.py) runs the natives through a fake JNIEnv.
"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = ["-I" + os.path.join(ROOT, "tests", "cpp", "jni_stub"), "-I" + os.path.join(ROOT, "include")]
SHIM = os.path.join(ROOT, "integration", "mhapb_jni.c")
LIBDIR = os.path.join(ROOT, "mhap_b200")


def build_shim(out):
    subprocess.check_call(["gcc", "-std=c11", "-O2", "-Wall", "-Wextra", "-Werror", "-fPIC", "-shared", *INC, SHIM, "-L" + LIBDIR, "-lmhap_b200",
                           "-Wl,-rpath," + LIBDIR, "-o", out])


def build_harness(out):
    subprocess.check_call(["gcc", "-std=c11", "-O1", "-Wall", "-Wextra", *INC, os.path.join(ROOT, "tests", "cpp", "jni_harness.c"), SHIM,
                           "-L" + LIBDIR, "-lmhap_b200", "-Wl,-rpath," + LIBDIR, "-o", out])


def test_shim_compiles_and_exports_every_native(tmp_path):
    so = str(tmp_path / "libmhapb_jni.so")
    build_shim(so)
    syms = subprocess.check_output(["nm", "-D", "--defined-only", so], text=True)
    exported = set(re.findall(r"Java_edu_umd_marbl_mhap_impl_MhapB200_(\w+)", syms))
    java = open(os.path.join(ROOT, "integration", "MhapB200.java")).read()
    declared = set(re.findall(r"static\s+native\s+[\w\[\]\.]+\s+(\w+)\s*\(", java))
    assert declared and declared == exported, (sorted(declared - exported), sorted(exported - declared))


def test_native_signatures_match_between_java_and_c():
    # argument counts: C has (JNIEnv*, jclass) + the Java parameters
    java = open(os.path.join(ROOT, "integration", "MhapB200.java")).read()
    c = open(SHIM).read()
    for name, params in re.findall(r"static\s+native\s+[\w\[\]\.]+\s+(\w+)\s*\(([^)]*)\)", java):
        n_java = len([x for x in params.split(",") if x.strip()])
        m = re.search(r"Java_edu_umd_marbl_mhap_impl_MhapB200_%s\s*\(([^)]*)\)" % name, c, re.S)
        assert m, name
        n_c = len([x for x in m.group(1).split(",") if x.strip()])
        assert n_c == n_java + 2, (name, n_java, n_c)


def test_glue_uses_only_natives_that_exist():
    java = open(os.path.join(ROOT, "integration", "MhapB200.java")).read()
    declared = set(re.findall(r"static\s+native\s+[\w\[\]\.]+\s+(\w+)\s*\(", java))
    glue = open(os.path.join(ROOT, "integration", "GpuMinHashSearch.java")).read()
    used = set(re.findall(r"MhapB200\.(\w+)\s*\(", glue))
    assert used <= declared, sorted(used - declared)
    # the seams of AbstractMatchSearch (impl/AbstractMatchSearch.java:119,121,201,203,312,314,340) are all overridden
    for sig in ("protected boolean addSequence(SequenceSketch", "public ArrayList<MatchResult> findMatches()",
                "protected List<MatchResult> findMatches(SequenceSketch", "public ArrayList<MatchResult> findMatches(final SequenceSketchStreamer",
                "public List<SequenceId> getStoredForwardSequenceIds()", "public SequenceSketch getStoredSequenceHash(SequenceId", "public int size()"):
        assert sig in glue, sig
    assert glue.count("{") == glue.count("}") and glue.count("(") == glue.count(")")


def test_harness_links(tmp_path):
    build_harness(str(tmp_path / "jni_harness"))
