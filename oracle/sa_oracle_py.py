"""ctypes binding of oracle/libsa_oracle.so (the CPU restatement, oracle/sa_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under segalign_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libsa_oracle.so"

SEGMENT_DTYPE = np.dtype([("ref_start", "<u4"), ("query_start", "<u4"), ("len", "<u4"), ("score", "<i4")])


class SaoShape(C.Structure):
    _fields_ = [("shape_pos", C.c_int * 32), ("transition_pos", C.c_int * 32),
                ("weight", C.c_int), ("span", C.c_int)]


class SaoParams(C.Structure):
    _fields_ = [("sub_mat", C.c_int * 64), ("xdrop", C.c_int), ("hspthresh", C.c_int),
                ("noentropy", C.c_int), ("seed_size", C.c_uint32), ("max_hits", C.c_uint32)]


class SaoTable(C.Structure):
    _fields_ = [("index", C.POINTER(C.c_uint32)), ("index_size", C.c_uint32),
                ("pos", C.POINTER(C.c_uint32)), ("num_pos", C.c_uint32)]


def _load() -> C.CDLL:
    if not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < (HERE / "sa_oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "libsa_oracle.so"], check=True, stdout=subprocess.DEVNULL)
    lib = C.CDLL(str(LIB_PATH))
    lib.sao_encode.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    lib.sao_encode.restype = None
    lib.sao_encode_rc.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    lib.sao_encode_rc.restype = None
    lib.sao_revcomp_ascii.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.sao_revcomp_ascii.restype = None
    lib.sao_shape_init.argtypes = [C.POINTER(SaoShape), C.c_char_p]
    lib.sao_kmer_at.argtypes = [C.POINTER(SaoShape), C.c_void_p, C.c_size_t]
    lib.sao_kmer_at.restype = C.c_uint32
    lib.sao_build_matrix.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int)]
    lib.sao_build_matrix.restype = None
    lib.sao_table_build.argtypes = [C.POINTER(SaoTable), C.POINTER(SaoShape), C.c_void_p, C.c_size_t,
                                    C.c_uint32, C.c_uint32]
    lib.sao_table_free.argtypes = [C.POINTER(SaoTable)]
    lib.sao_table_free.restype = None
    lib.sao_chunk_seeds.argtypes = [C.POINTER(SaoShape), C.c_int, C.c_void_p, C.c_size_t, C.c_uint32,
                                    C.c_uint32, C.c_void_p]
    lib.sao_chunk_seeds.restype = C.c_size_t
    lib.sao_extend_hit.argtypes = [C.POINTER(SaoParams), C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                   C.c_uint32, C.c_uint32, C.c_void_p]
    lib.sao_seed_and_filter.argtypes = [C.POINTER(SaoParams), C.POINTER(SaoTable), C.c_void_p, C.c_uint32,
                                        C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                        C.POINTER(C.c_size_t)]
    lib.sao_seed_and_filter.restype = C.c_void_p
    lib.sao_iteration_plan.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    lib.sao_sort_dedupe.argtypes = [C.c_void_p, C.c_size_t]
    lib.sao_sort_dedupe.restype = C.c_size_t
    lib.sao_rm_revcomp_codes.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    lib.sao_rm_revcomp_codes.restype = None
    lib.sao_rm_sort_dedupe.argtypes = [C.c_void_p, C.c_size_t]
    lib.sao_rm_sort_dedupe.restype = C.c_size_t
    lib.sao_rm_seed_and_filter.argtypes = [C.POINTER(SaoParams), C.POINTER(SaoTable), C.c_void_p, C.c_void_p, C.c_uint32,
                                           C.c_void_p, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.POINTER(C.c_size_t)]
    lib.sao_rm_seed_and_filter.restype = C.c_void_p
    lib.sao_free.argtypes = [C.c_void_p]
    lib.sao_free.restype = None
    return lib


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _u8(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint8)


def encode(seq) -> np.ndarray:
    seq = _u8(seq)
    out = np.empty(seq.size, dtype=np.uint8)
    lib().sao_encode(seq.ctypes.data, seq.size, out.ctypes.data)
    return out


def encode_rc(seq):
    seq = _u8(seq)
    fwd = np.empty(seq.size, dtype=np.uint8)
    rc = np.empty(seq.size, dtype=np.uint8)
    lib().sao_encode_rc(seq.ctypes.data, seq.size, fwd.ctypes.data, rc.ctypes.data)
    return fwd, rc


def revcomp_ascii(seq) -> np.ndarray:
    seq = _u8(seq)
    out = np.empty(seq.size, dtype=np.uint8)
    lib().sao_revcomp_ascii(out.ctypes.data, seq.ctypes.data, seq.size)
    return out


def build_matrix(ambiguous: str, xdrop: int) -> np.ndarray:
    m = (C.c_int * 64)()
    lib().sao_build_matrix((ambiguous or "").encode(), xdrop, m)
    return np.array(m, dtype=np.int32)


class Shape:
    def __init__(self, seed_shape: str):
        self.c = SaoShape()
        self.weight = lib().sao_shape_init(C.byref(self.c), seed_shape.encode())
        self.span = self.c.span

    def kmer_at(self, seq: np.ndarray, pos: int) -> int:
        return lib().sao_kmer_at(C.byref(self.c), seq.ctypes.data, pos)

    def chunk_seeds(self, seq: np.ndarray, j0: int, j1: int, transition: bool) -> np.ndarray:
        """seq must extend at least span-1 bytes past j1 (the DRAM arena is zero-filled)."""
        out = np.empty(max(1, (j1 - j0) * (1 + self.weight)), dtype=np.uint64)
        n = lib().sao_chunk_seeds(C.byref(self.c), int(transition), seq.ctypes.data, 0, j0, j1,
                                  out.ctypes.data)
        return out[:n].copy()


class Table:
    def __init__(self, shape: Shape, ref: np.ndarray, ref_length: int, step: int = 1):
        self.c = SaoTable()
        self._ref = _u8(ref)
        rc = lib().sao_table_build(C.byref(self.c), C.byref(shape.c), self._ref.ctypes.data, 0,
                                   ref_length, step)
        if rc != 0:
            raise MemoryError("sao_table_build failed")

    @property
    def index(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.c.index, shape=(self.c.index_size,))

    @property
    def pos(self) -> np.ndarray:
        if self.c.num_pos == 0:
            return np.empty(0, dtype=np.uint32)
        return np.ctypeslib.as_array(self.c.pos, shape=(self.c.num_pos,))

    def __del__(self):
        try:
            lib().sao_table_free(C.byref(self.c))
        except Exception:
            pass


def make_params(sub_mat, xdrop, hspthresh, noentropy, seed_size, max_hits) -> SaoParams:
    p = SaoParams()
    for i, v in enumerate(np.asarray(sub_mat, dtype=np.int32).reshape(64)):
        p.sub_mat[i] = int(v)
    p.xdrop, p.hspthresh, p.noentropy = int(xdrop), int(hspthresh), int(bool(noentropy))
    p.seed_size, p.max_hits = int(seed_size), int(max_hits)
    return p


def extend_hit(params: SaoParams, ref_enc, qry_enc, r0: int, q0: int):
    ref_enc, qry_enc = _u8(ref_enc), _u8(qry_enc)
    seg = np.zeros(1, dtype=SEGMENT_DTYPE)
    ok = lib().sao_extend_hit(C.byref(params), ref_enc.ctypes.data, ref_enc.size, qry_enc.ctypes.data,
                              qry_enc.size, r0, q0, seg.ctypes.data)
    return bool(ok), seg[0]


def seed_and_filter(params: SaoParams, table: Table, ref_enc, qry_enc, seeds) -> np.ndarray:
    ref_enc, qry_enc = _u8(ref_enc), _u8(qry_enc)
    seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
    n = C.c_size_t()
    p = lib().sao_seed_and_filter(C.byref(params), C.byref(table.c), ref_enc.ctypes.data, ref_enc.size,
                                  qry_enc.ctypes.data, qry_enc.size, seeds.ctypes.data, seeds.size,
                                  C.byref(n))
    out = np.empty(n.value, dtype=SEGMENT_DTYPE)
    C.memmove(out.ctypes.data, p, n.value * 16)
    lib().sao_free(p)
    return out


def iteration_plan(prefix, max_hits: int) -> np.ndarray:
    prefix = np.ascontiguousarray(prefix, dtype=np.uint32)
    nh = int(prefix[-1]) if prefix.size else 0
    lim = np.zeros(nh // max(1, max_hits) + 2, dtype=np.uint32)
    n = lib().sao_iteration_plan(prefix.ctypes.data, prefix.size, max_hits, lim.ctypes.data)
    return lim[:n].copy()


def sort_dedupe(segs) -> np.ndarray:
    a = np.ascontiguousarray(segs, dtype=SEGMENT_DTYPE).copy()
    n = lib().sao_sort_dedupe(a.ctypes.data, a.size)
    return a[:n]


# ---- repeat-masker variant (repeat_masker_src/seed_filter.cu)
def rm_revcomp_codes(enc) -> np.ndarray:
    enc = _u8(enc)
    out = np.empty(enc.size, dtype=np.uint8)
    lib().sao_rm_revcomp_codes(enc.ctypes.data, enc.size, out.ctypes.data)
    return out


def rm_sort_dedupe(segs) -> np.ndarray:
    a = np.ascontiguousarray(segs, dtype=SEGMENT_DTYPE).copy()
    n = lib().sao_rm_sort_dedupe(a.ctypes.data, a.size)
    return a[:n]


def rm_seed_and_filter(params: SaoParams, table: Table, seq_enc, seq_rc_enc, seeds, rev: bool, ref_start: int,
                       ref_end: int) -> np.ndarray:
    seq_enc, seq_rc_enc = _u8(seq_enc), _u8(seq_rc_enc)
    seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
    n = C.c_size_t()
    p = lib().sao_rm_seed_and_filter(C.byref(params), C.byref(table.c), seq_enc.ctypes.data, seq_rc_enc.ctypes.data,
                                     seq_enc.size, seeds.ctypes.data, seeds.size, int(rev), ref_start, ref_end, C.byref(n))
    out = np.empty(n.value, dtype=SEGMENT_DTYPE)
    C.memmove(out.ctypes.data, p, n.value * 16)
    lib().sao_free(p)
    return out
