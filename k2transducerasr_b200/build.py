"""In-tree build of libk2b200.so with explicit nvcc for sm_100a (no JIT cache: the built .so travels with
the repo snapshot to the GPU box).  `python -m k2transducerasr_b200.build [--force]`."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libk2b200.so"
STAMP = PKG / "csrc" / ".build_stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; libk2b200.so cannot be built (there is no CPU fallback)")
    return cand


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) +
                    [PKG.parent / "include" / "k2b200.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Idempotent and safe under torchrun: the ranks of one box serialise on a file lock, the first one compiles, the others find
    the stamp up to date when they get the lock."""
    import fcntl
    digest = _digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB
    with open(PKG / ".build_lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
                return LIB
            return _build_locked(digest, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(digest: str, verbose: bool) -> Path:
    nvcc = _nvcc()
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    objs = []
    procs = []
    for src in sources():
        obj = objdir / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-Xptxas", "-v", "-c", str(src), "-o", str(obj)]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src.name}\n{out}")
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {src.name}")
    (objdir / "ptxas.log").write_text("\n".join(log))
    if verbose:
        print("\n".join(log))
    tmp = LIB.with_suffix(".so.tmp")
    cmd = [nvcc, "-shared", "-o", str(tmp), *map(str, objs), "-Xlinker", "--exclude-libs,ALL", "-cudart", "static", "-ldl"]
    subprocess.run(cmd, check=True)
    os.replace(tmp, LIB)              # a process that already mapped the old file keeps it
    STAMP.write_text(digest)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
