"""-m gpu: every native of integration/mhapb_jni.c driven through a fake JNIEnv on the GPU (tests/cpp/jni_harness.c) and
compared with direct C-ABI calls on the same reads: store, self search, store-vs-query (FASTA and .dat routes), the
per-sequence seam, stored-sketch read-back, the -f filter, and the reference's error text as the exception message."""
import subprocess

import pytest

from tests.test_jni_shim import build_harness

pytestmark = pytest.mark.gpu


def test_jni_shim_natives_against_the_c_abi(tmp_path):
    exe = str(tmp_path / "jni_harness")
    build_harness(exe)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "JNI_HARNESS_OK" in r.stdout
