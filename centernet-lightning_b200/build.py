"""Build the C-ABI shared library (libcnl_b200.so) in-tree with nvcc for sm_100a.

In-tree on purpose: the built .so travels with the repo snapshot to the GPU box, a JIT cache would not."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcnl_b200.so")
SOURCES = ["cnl_decode.cu", "cnl_conv.cu", "cnl_io.cu", "cnl_track.cu"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the cnl_b200 CUDA library cannot be built")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.endswith(".o")]
    deps.append(os.path.join(HERE, "..", "include", "cnl_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, defines=(), lib_out: str = LIB) -> str:
    """Compile csrc/*.cu -> libcnl_b200.so.  Returns the library path.  `defines` / `lib_out` build an A/B variant of the same
    library next to it (loaded through CNL_LIB for timing experiments; never a fallback)."""
    variant = lib_out != LIB
    if not force and not variant and not _stale():
        return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    tag = "" if not variant else "." + os.path.splitext(os.path.basename(lib_out))[0]
    for s in SOURCES:
        obj = os.path.join(CSRC, s.replace(".cu", tag + ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", os.path.join(CSRC, s), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
        if verbose:
            print(out, file=sys.stderr)
    cmd = [nvcc, "-shared", "-o", lib_out, *objs, "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return lib_out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
