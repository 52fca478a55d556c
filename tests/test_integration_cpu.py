"""Drop-in checks that need no GPU: the reference's own, unmodified cuDNN-facing sources and the
warp-ctc call-site shape compile against the headers this repo ships (integration/kaldi)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present on this box")
def test_reference_sources_compile_against_dropin_headers():
    out = subprocess.run([os.path.join(ROOT, "integration", "kaldi", "check_compile.sh")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "OK: reference sources compile" in out.stdout


def test_ctc_callsite_shape_compiles_without_reference():
    src = os.path.join(ROOT, "integration", "kaldi", "ctc_callsite_check.cc")
    out = subprocess.run(["/usr/bin/g++", "-std=c++11", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "include"),
                          src], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
