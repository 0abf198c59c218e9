"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/*.h declares.
No compute calls here (no GPU in the CPU container)."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "segalign_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sa_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_reference_boundary(built):
    syms = declared_symbols()
    for s in ("sa_initialize_interface", "sa_initialize_processor", "sa_send_ref", "sa_clear_ref",
              "sa_generate_seed_pos_table", "sa_send_query", "sa_clear_query", "sa_seed_and_filter",
              "sa_shutdown_processor"):
        assert s in syms


def test_library_exports_every_declared_symbol(built):
    from segalign_b200 import backend
    lib = ctypes.CDLL(str(backend.LIB_PATH))
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/segalign_b200.h but not exported: {missing}"
    assert sorted(backend.ABI_SYMBOLS) == declared_symbols()


def test_library_is_sm100a_only(built):
    from segalign_b200 import backend
    out = subprocess.run(["cuobjdump", "-lelf", str(backend.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_segment_layout_is_16_bytes(built):
    from segalign_b200 import backend
    assert backend.SEGMENT_DTYPE.itemsize == 16  # src/graph.h:25-30


def test_no_gpu_fails_loudly(built):
    """Without a device the product path must error (reference: exit(1), seed_filter_interface.cu:54-57)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from segalign_b200.backend import Backend, BackendError
    be = Backend()
    with pytest.raises(BackendError) as ei:
        be.InitializeInterface(1)
    assert ei.value.code == -1
    with pytest.raises(BackendError):
        be.SeedAndFilter([1, 2, 3], False, 0)


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/."""
    for p in (ROOT / "segalign_b200").rglob("*"):
        if p.name == "build.py":  # compiles the checker (make -C oracle) but never loads it
            continue
        if p.suffix in (".py", ".cu", ".cuh", ".cpp", ".h") and p.is_file():
            text = p.read_text()
            for pat in (r'#\s*include\s*[<"][^>"]*oracle', r"^\s*(from|import)\s+oracle\b", r"libsa_oracle", r"dlopen"):
                assert not re.search(pat, text, flags=re.M), (p, pat)


def test_sass_has_no_misencoded_async_copies(built):
    """ptxas 12.9 mis-encodes `cp.async ... L2::cache_hint` copies whose shared address is a previous copy's
    plus an immediate: the LDGSTS then names an unset, odd uniform register as its descriptor
    (`[R41+UR0+0x200], desc[UR1]`) and the launch dies with "illegal instruction" (seen on a B200, round 2).
    Every 64-bit descriptor operand in the library's SASS must be an even uniform register."""
    import re
    import subprocess
    from segalign_b200.backend import LIB_PATH
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB_PATH)], capture_output=True, text=True, check=True).stdout
    ops = [l for l in sass.splitlines() if re.search(r"\b(LDGSTS|LDG|STG|ATOMG|REDG)\b", l) and "desc[" in l]
    assert len(ops) > 100
    bad = [l.strip() for l in ops if int(re.search(r"desc\[UR(\d+)\]", l).group(1)) % 2]
    assert not bad, bad[:5]
