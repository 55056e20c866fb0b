"""In-tree build of the sm_100a C-ABI library (``libb200_attn_mlp.so``).

nvcc cross-compiles without a GPU, so this runs in the authoring container as well as on a B200 box. The
library is built next to the sources (git-ignored, but it travels to the GPU box with the repo snapshot).
Flags: exactly ``-gencode arch=compute_100a,code=sm_100a`` — the shorthand ``-arch=sm_100a`` adds a generic
compute_100 PTX pass in which ptxas rejects every tcgen05 instruction.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_NAME = "libb200_attn_mlp.so"
LIB_PATH = PKG_DIR / LIB_NAME
SOURCES = ["host_common.cu", "fa_fwd.cu", "fa_decode.cu", "gemm_mlp.cu", "ln_kernels.cu", "tp_allreduce.cu"]
HEADERS = ["common.cuh", "host_common.h", "../../include/b200_attn_mlp.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build the sm_100a extension")


def _existing_sources() -> list[Path]:
    return [CSRC / s for s in SOURCES if (CSRC / s).exists()]


STAMP_PATH = PKG_DIR / "build" / "source_stamp.txt"


def source_stamp() -> str:
    """SHA-256 over every source, header and the compiler flags: what the library on disk must have been built from."""
    import hashlib

    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS + os.environ.get("B200_EXTRA_NVCC_FLAGS", "").split()).encode())
    for d in _existing_sources() + [(CSRC / x).resolve() for x in HEADERS]:
        if d.exists():
            h.update(d.name.encode())
            h.update(d.read_bytes())
    return h.hexdigest()


def needs_build() -> bool:
    """True unless the library exists AND was built from exactly the current sources (content hash, not mtime: a
    checkout or a snapshot copy resets mtimes, and a prebuilt .so must never mask a source change)."""
    if not LIB_PATH.exists() or not STAMP_PATH.exists():
        return True
    return STAMP_PATH.read_text().strip() != source_stamp()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a and link the shared library. Returns its path."""
    force = force or os.environ.get("B200_FORCE_REBUILD", "") not in ("", "0")
    if not force and not needs_build():
        return LIB_PATH
    stamp = source_stamp()
    nvcc = _nvcc()
    obj_dir = PKG_DIR / "build"
    obj_dir.mkdir(exist_ok=True)
    procs = []
    objs = []
    for src in _existing_sources():
        obj = obj_dir / (src.stem + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("B200_EXTRA_NVCC_FLAGS", "").split(), "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, proc in procs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name}:\n{out}")
        if verbose and out:
            print(out)
    tmp = LIB_PATH.with_suffix(".so.tmp")
    cmd = [nvcc, "-shared", "-o", str(tmp), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}")
    os.replace(tmp, LIB_PATH)
    STAMP_PATH.write_text(stamp + "\n")
    return LIB_PATH


if __name__ == "__main__":
    import sys

    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
