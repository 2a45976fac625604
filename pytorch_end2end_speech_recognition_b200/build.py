"""In-tree build of the CUDA library (sm_100a only) with plain nvcc.

The C-ABI library has no PyTorch dependency, so it is compiled directly:
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -shared -Xcompiler -fPIC ...
The result lives next to the sources (``lib/libb200ctc.so``): it is git-ignored but
travels to the GPU box with the repo snapshot.
"""

import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.environ.get("B200CTC_LIB") or os.path.join(LIB_DIR, "libb200ctc.so")   # B200CTC_LIB: developer variants

SOURCES = ["api.cu", "plan.cu", "softmax_rows.cu", "lattice.cu", "greedy.cu", "beam.cu", "eval.cu"]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def _nvcc():
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cand = os.path.join(cuda_home, "bin", "nvcc")
    return cand if os.path.exists(cand) else "nvcc"


def _source_hash(deps, flags):
    """Content hash of everything the library is built from (sources, header, flags): file times do not
    survive the snapshot copy to the GPU box, contents do."""
    import hashlib
    h = hashlib.sha1(" ".join(flags).encode())
    for d in sorted(deps):
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale(target, deps, flags=()):
    if not os.path.exists(target):
        return True
    try:
        with open(target + ".srchash") as f:
            return f.read().strip() != _source_hash(deps, flags)
    except OSError:
        return True


def build_library(force=False, verbose=False, extra_flags=(), lib_path=None):
    """Compile csrc/*.cu into lib/libb200ctc.so (no-op when up to date).  ``extra_flags`` /
    ``lib_path`` build a developer variant next to it (tools/trace_lattice.py: -DB200CTC_TRACE)."""
    os.makedirs(LIB_DIR, exist_ok=True)
    if lib_path is not None:
        return _build_variant(list(extra_flags), lib_path, verbose)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(REPO_DIR, "include", "b200ctc.h")]
    if not force and not _stale(LIB_PATH, deps, NVCC_FLAGS):
        return LIB_PATH
    objs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        cmd = [_nvcc()] + NVCC_FLAGS + ["-I", os.path.join(REPO_DIR, "include"), "-I", CSRC,
                                        "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
        objs.append(obj)
    cmd = [_nvcc(), "-shared", "-o", LIB_PATH] + objs + ["-Xcompiler", "-fPIC"]
    subprocess.run(cmd, check=True)
    with open(LIB_PATH + ".srchash", "w") as f:
        f.write(_source_hash(deps, NVCC_FLAGS))
    return LIB_PATH


def _build_variant(extra_flags, lib_path, verbose):
    objs = []
    tag = os.path.basename(lib_path).replace(".so", "")
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, "%s_%s" % (tag, src.replace(".cu", ".o")))
        cmd = [_nvcc()] + NVCC_FLAGS + extra_flags + ["-I", os.path.join(REPO_DIR, "include"), "-I", CSRC,
                                                      "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
        objs.append(obj)
    subprocess.run([_nvcc(), "-shared", "-o", lib_path] + objs + ["-Xcompiler", "-fPIC"], check=True)
    return lib_path


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
