"""In-tree build of the CUDA backend (sm_100a only) and of the test-only oracle.

`nvcc` cross-compiles without a GPU, so this runs in the CPU container as the "does it build"
check and the resulting .so files travel to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "segalign_b200" / "csrc"
LIB = ROOT / "segalign_b200" / "libsegalign_b200.so"
CLI = ROOT / "segalign_b200" / "segalign_b200_cli"
SHIM_OBJ = ROOT / "segalign_b200" / "libsegalign_b200_shim.a"
ORACLE_DIR = ROOT / "oracle"
ORACLE_LIB = ORACLE_DIR / "libsa_oracle.so"
REFERENCE = Path(os.environ.get("SEGALIGN_REFERENCE", "/root/reference"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA backend cannot be built")


def _newer(target: Path, sources) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(s).stat().st_mtime <= t for s in sources)


def build_backend(force: bool = False, verbose: bool = False) -> Path:
    sources = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.inc")) + sorted(CSRC.glob("*.cpp")) + [ROOT / "include" / "segalign_b200.h"]
    if not force and _newer(LIB, sources):
        return LIB
    extra = os.environ.get("SEGALIGN_B200_NVCC_EXTRA", "").split()  # e.g. -DSA_SCR_THREADS=288 (tuning experiments)
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-o", str(LIB), str(CSRC / "sa_backend.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True, cwd=str(CSRC))
    # command line front end of the whole-genome driver (host only; links the library by rpath)
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(CLI), str(CSRC / "segalign_main.cpp"),
                    f"-L{LIB.parent}", "-lsegalign_b200", f"-Wl,-rpath,{LIB.parent}"], check=True)
    return LIB


def build_oracle(force: bool = False) -> Path:
    srcs = [ORACLE_DIR / "sa_oracle.c", ORACLE_DIR / "sa_oracle.h"]
    if force or not _newer(ORACLE_LIB, srcs):
        subprocess.run(["make", "-C", str(ORACLE_DIR), "libsa_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return ORACLE_LIB


def build_reference_oracle(force: bool = False) -> bool:
    """oracle/_ref/{oracle_runner,new_runner,lastz}: only where the reference tree exists."""
    if not REFERENCE.exists():
        return False
    targets = ["ref"]
    if force:
        subprocess.run(["make", "-C", str(ORACLE_DIR), "clean"], check=True, stdout=subprocess.DEVNULL)
    subprocess.run(["make", "-C", str(ORACLE_DIR), f"REF={REFERENCE}", *targets], check=True,
                   stdout=subprocess.DEVNULL)
    return True


def build_all(force: bool = False) -> None:
    build_backend(force)
    build_oracle(force)
    build_reference_oracle(False)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built:", LIB, ORACLE_LIB)
