"""In-tree build of ``libswe_gnn_b200.so`` with nvcc for sm_100a (cross-compiles without a GPU).

    python -m mswe_gnn_b200.build            # or __graft_entry__.build()

The shared object is written next to the sources (``mswe-gnn_b200/csrc/``) so that it travels
with the repository snapshot to the GPU box; it is git-ignored.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_NAME = "libswe_gnn_b200.so"
LIB_PATH = os.path.join(CSRC, LIB_NAME)
SOURCES = ["swe_plan.cu", "swe_forward.cu", "swe_backward.cu", "swe_gate_tc.cu", "swe_gate_tc16.cu", "swe_hop_tc.cu", "swe_hop_tc16.cu", "swe_rowmlp_tc.cu", "swe_rowlin_tc16.cu", "swe_rowmlp_tc16.cu", "swe_train_tc.cu", "swe_halo.cu", "swe_train_step.cu", "swe_dataset.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build " + LIB_NAME)


def _stamp(paths) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(HERE, "..", "include", "swe_gnn_b200.h"))
    stamp_file = os.path.join(CSRC, ".build_stamp")
    stamp = _stamp(deps)
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp_file) \
            and open(stamp_file).read().strip() == stamp:
        return LIB_PATH
    nvcc = _nvcc()
    objs = [os.path.join(CSRC, os.path.splitext(os.path.basename(s))[0] + ".o") for s in srcs]
    logs = []

    def compile_one(pair):
        src, obj = pair
        r = subprocess.run([nvcc, *NVCC_FLAGS, "-c", src, "-o", obj], capture_output=True, text=True)
        return src, r

    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        for src, r in ex.map(compile_one, zip(srcs, objs)):
            logs.append(f"== {os.path.basename(src)}\n{r.stderr}")
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    r = subprocess.run([nvcc, "-shared", "-o", LIB_PATH, *objs, "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(os.path.join(CSRC, "ptxas_info.log"), "w") as f:
        f.write("\n".join(logs))
    with open(stamp_file, "w") as f:
        f.write(stamp)
    if verbose:
        print("\n".join(logs))
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
