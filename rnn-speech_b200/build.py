"""In-tree build of librnnspeech_b200.so (nvcc, sm_100a only).

    python rnn-speech_b200/build.py [--force]

The shared library is written next to this file so that it travels with the
gpurun snapshot; it is git-ignored (*.so).
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librnnspeech_b200.so")
LIB_DIAG = os.path.join(HERE, "librnnspeech_b200_diag.so")      # + the diagnostic hooks (-DRS_DIAG), for tests only
DIAG_SOURCES = ("tc_selftest.cu", "gemm_tc.cu")
STAMP = os.path.join(HERE, "build", "stamp.txt")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(f.encode() + b"\0" + fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(LIB_DIAG) and os.path.exists(STAMP) \
            and open(STAMP).read().strip() == digest:
        return LIB
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(HERE, "build", os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    diag_objs = {}
    for name in DIAG_SOURCES:
        obj = os.path.join(HERE, "build", "diag_" + name[:-3] + ".o")
        diag_objs[name[:-3] + ".o"] = obj
        cmd = [NVCC] + FLAGS + ["-DRS_DIAG", "-c", os.path.join(CSRC, name), "-o", obj]
        procs.append((name + " (diag)", subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append("== %s\n%s" % (os.path.basename(src), out))
        if p.returncode != 0:
            failed = True
    with open(os.path.join(HERE, "build", "nvcc.log"), "w") as fh:
        fh.write("\n".join(log))
    if failed or verbose:
        sys.stderr.write("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed, see rnn-speech_b200/build/nvcc.log")
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lcudart", "-ldl"]
    subprocess.check_call(cmd)
    # the diagnostic library: the same objects, with the two translation units that carry hooks rebuilt with -DRS_DIAG
    dobjs = [diag_objs.get(os.path.basename(o), o) for o in objs]
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_DIAG] + dobjs + ["-lcudart", "-ldl"]
    subprocess.check_call(cmd)
    with open(STAMP, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
