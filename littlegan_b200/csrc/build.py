"""Build liblittlegan_b200.so (sm_100a only) in-tree with nvcc.

    python -m littlegan_b200.csrc.build [--force]

The library is a plain C-ABI shared object (include/littlegan_b200.h); it does not link
against libtorch.  It is git-ignored but travels with the repo snapshot to the GPU box.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SOURCES = ["api.cu", "simt_gemm.cu", "elementwise.cu", "fid.cu", "tc_conv.cu", "tc_wgrad.cu", "tc_deconv_small.cu", "tc_conv_cin3.cu", "tc_dgrad4.cu", "heads.cu", "tc_rowconv.cu", "tc_rowdeconv.cu", "tc_rowdgrad.cu", "tc_rowwgrad.cu", "augment.cu", "inception.cu", "tc_convbn.cu"]
HEADERS = ["common.cuh", "internal.h", "tc_ptx.cuh", "tc_host.cuh", "norm_bwd.cuh", os.path.join(ROOT, "include", "littlegan_b200.h")]
LIB = os.path.join(os.path.dirname(HERE), "liblittlegan_b200.so")
STAMP = LIB + ".stamp"
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-I", os.path.join(ROOT, "include"), "-I", HERE,
]


def _digest():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(HERE, f) if not os.path.isabs(f) else f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for s in SOURCES:
        o = os.path.join(objdir, s.replace(".cu", ".o"))
        objs.append(o)
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(HERE, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write("== %s ==\n%s\n" % (s, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building liblittlegan_b200.so")
    subprocess.check_call([nvcc, "-shared", "-o", LIB] + objs + ["-lcudart"])
    with open(STAMP, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
