"""Locates (and, in the container that has /root/reference, builds) the binaries that oracle/ref/Makefile
links from the reference's own unmodified sources: oracle/_ref/ref_component_harness and
oracle/_ref/ref_ctc_train (= src/ctcbin/nnet2-ctc-train-simple.cc).  Test infrastructure only."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
HARNESS = os.path.join(REFDIR, "ref_component_harness")
TRAIN = os.path.join(REFDIR, "ref_ctc_train")
NNET3 = os.path.join(REFDIR, "ref_nnet3_harness")
CPULIB = os.path.join(REFDIR, "libkaldi_ref_cpu.so")


def ensure_built(target="all"):
    """Returns True when the binaries exist; builds them when the reference tree is present."""
    want = {"gpu": [HARNESS, TRAIN, NNET3], "cpu": [CPULIB], "all": [HARNESS, TRAIN, NNET3, CPULIB]}[target]
    if all(os.path.exists(p) for p in want):
        return True
    if not os.path.isdir("/root/reference/src"):
        return False
    subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "oracle", "ref"), target])
    return all(os.path.exists(p) for p in want)


def run(binary, *args, env=None, timeout=600):
    e = dict(os.environ)
    if env:
        e.update(env)
    r = subprocess.run([binary] + [str(a) for a in args], capture_output=True, text=True, env=e, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError("%s %s failed (%d):\n%s\n%s" % (os.path.basename(binary), " ".join(map(str, args)),
                                                          r.returncode, r.stdout[-3000:], r.stderr[-6000:]))
    return r
