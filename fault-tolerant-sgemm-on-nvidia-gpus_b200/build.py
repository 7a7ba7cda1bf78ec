"""In-tree build of libftsgemm.so and the ft_sgemm driver for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
ROOT = HERE.parent
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++"]

LIB = HERE / "libftsgemm.so"
CLI = HERE / "ft_sgemm"


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def build(force: bool = False, verbose: bool = False) -> None:
    lib_src = [CSRC / "ftsgemm.cu"]
    lib_dep = lib_src + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [ROOT / "include" / "ftsgemm.h"]
    if force or _stale(LIB, lib_dep):
        cmd = [NVCC, *ARCH, *COMMON, "-shared", "-Xptxas", "-v", *map(str, lib_src), "-o", str(LIB), "-lcublas"]
        out = subprocess.run(cmd, capture_output=True, text=True)
        (HERE / "build_ptxas.log").write_text(out.stderr)
        if verbose or out.returncode != 0:
            print(out.stdout, out.stderr, file=sys.stderr)
        if out.returncode != 0:
            raise RuntimeError("nvcc failed building libftsgemm.so")
    cli_src = CSRC / "ft_sgemm_main.cu"
    if cli_src.exists() and (force or _stale(CLI, [cli_src, LIB])):
        cmd = [NVCC, *ARCH, "-O3", "-std=c++17", "-ccbin", "/usr/bin/g++", str(cli_src), "-o", str(CLI),
               f"-L{HERE}", "-lftsgemm", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"]
        out = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or out.returncode != 0:
            print(out.stdout, out.stderr, file=sys.stderr)
        if out.returncode != 0:
            raise RuntimeError("nvcc failed building ft_sgemm")


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
