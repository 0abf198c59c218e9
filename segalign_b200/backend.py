"""ctypes binding of the C ABI (include/segalign_b200.h).

The method names mirror the reference backend boundary (common/seed_filter_interface.h:3-11,
src/seed_filter.h:4-14, common/ntcoding.h:9) so tests read like calls into SegAlign itself.
There is no fallback: if libsegalign_b200.so is missing or a CUDA call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

LIB_PATH = Path(__file__).resolve().parent / "libsegalign_b200.so"

SEGMENT_DTYPE = np.dtype([("ref_start", "<u4"), ("query_start", "<u4"), ("len", "<u4"), ("score", "<i4")])

# every symbol include/segalign_b200.h declares
ABI_SYMBOLS = [
    "sa_last_error", "sa_initialize_interface", "sa_initialize_interface_at",
    "sa_initialize_processor", "sa_set_max_hits", "sa_get_max_hits", "sa_set_filter_kernel", "sa_set_seed_shape",
    "sa_send_ref", "sa_generate_seed_pos_table", "sa_clear_ref", "sa_send_query",
    "sa_clear_query", "sa_seed_and_filter", "sa_release_result", "sa_seed_and_filter_range",
    "sa_shutdown_processor", "sa_debug_get_table", "sa_debug_get_encoded", "sa_get_stats",
    "sa_reset_stats", "sa_set_profiling", "sa_version", "sa_host_chunk_seeds", "sa_write_segments",
    "sa_pipeline_run", "sa_pipeline_plan", "sa_build_matrix", "sa_get_gpu_calls",
    "sa_rm_send_query", "sa_rm_clear_query", "sa_rm_seed_and_filter", "sa_rm_seed_and_filter_range",
]


class SaStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("calls", "seeds", "hits", "survivors", "anchors_pre_dedupe", "hsps", "ext_cells")] + \
               [(n, C.c_double) for n in
                ("ms_h2d", "ms_count_scan", "ms_lookup", "ms_prefilter", "ms_extend", "ms_sort",
                 "ms_d2h", "ms_ref_encode", "ms_table_build", "ms_query_encode")] + \
               [("launches", C.c_uint64), ("walked", C.c_uint64), ("h2d_bytes", C.c_uint64),
                ("merge_calls", C.c_uint64), ("merge_dropped", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class SaPipelineConfig(C.Structure):
    _fields_ = [(n, C.c_char_p) for n in
                ("ref_fasta", "query_fasta", "out_dir", "data_folder", "seed_shape", "strand", "ambiguous",
                 "output_format", "scoring_file")] + \
               [("sub_mat", C.POINTER(C.c_int))] + \
               [(n, C.c_int) for n in ("transition", "noentropy", "gapped", "notrivial", "xdrop", "ydrop",
                                       "hspthresh", "gappedthresh")] + \
               [(n, C.c_uint32) for n in ("step", "wga_chunk", "lastz_interval")] + \
               [("seq_block_size", C.c_uint64), ("num_gpu", C.c_int), ("num_threads", C.c_int)]


class SaPipelineReport(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("ref_blocks", "query_blocks", "intervals", "calls", "seeds", "hits",
                                          "hsps", "segment_files")] + \
               [(n, C.c_double) for n in ("seconds", "ms_ref_upload", "ms_table_build", "ms_query_upload",
                                          "seconds_read_input", "seconds_device_init", "seconds_align")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class SaChromTable(C.Structure):
    _fields_ = [("names", C.POINTER(C.c_char_p)), ("starts", C.POINTER(C.c_uint64)),
                ("lens", C.POINTER(C.c_uint32)), ("count", C.c_uint32)]

    @classmethod
    def build(cls, names, starts, lens):
        n = len(names)
        t = cls()
        t._names = (C.c_char_p * n)(*[s.encode() for s in names])
        t._starts = (C.c_uint64 * n)(*[int(x) for x in starts])
        t._lens = (C.c_uint32 * n)(*[int(x) for x in lens])
        t.names, t.starts, t.lens, t.count = t._names, t._starts, t._lens, n
        return t


class BackendError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"segalign_b200 error {code}: {msg}")
        self.code = code


def load_library(path: Path | None = None) -> C.CDLL:
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise FileNotFoundError(
            f"{p} not found: build it with `python -m segalign_b200.build` (nvcc, sm_100a). "
            "There is no CPU fallback.")
    lib = C.CDLL(str(p))
    lib.sa_last_error.restype = C.c_char_p
    lib.sa_version.restype = C.c_char_p
    lib.sa_get_max_hits.restype = C.c_uint32
    lib.sa_initialize_interface.argtypes = [C.c_int]
    lib.sa_initialize_interface_at.argtypes = [C.c_int, C.c_int]
    lib.sa_initialize_processor.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.POINTER(C.c_int),
                                            C.c_int, C.c_int, C.c_int]
    lib.sa_set_max_hits.argtypes = [C.c_uint32]
    lib.sa_set_filter_kernel.argtypes = [C.c_int]
    lib.sa_build_matrix.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int)]
    lib.sa_pipeline_run.argtypes = [C.c_void_p, C.c_void_p]
    lib.sa_pipeline_plan.argtypes = [C.c_void_p, C.c_void_p]
    lib.sa_set_seed_shape.argtypes = [C.c_char_p]
    lib.sa_send_ref.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32]
    lib.sa_generate_seed_pos_table.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_int]
    lib.sa_send_query.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32]
    lib.sa_clear_query.argtypes = [C.c_uint32]
    lib.sa_seed_and_filter.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_uint32,
                                       C.POINTER(C.c_void_p), C.POINTER(C.c_uint32)]
    lib.sa_release_result.argtypes = [C.c_void_p]
    lib.sa_release_result.restype = None
    lib.sa_seed_and_filter_range.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_uint32,
                                             C.POINTER(C.c_void_p), C.POINTER(C.c_uint32),
                                             C.POINTER(C.c_uint32)]
    lib.sa_debug_get_table.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_void_p, C.c_void_p]
    lib.sa_debug_get_encoded.argtypes = [C.c_int, C.c_uint32, C.c_void_p, C.c_uint32]
    lib.sa_get_stats.argtypes = [C.POINTER(SaStats)]
    lib.sa_set_profiling.argtypes = [C.c_int]
    lib.sa_get_gpu_calls.argtypes = [C.POINTER(C.c_uint64), C.c_int]
    lib.sa_rm_seed_and_filter.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32,
                                          C.POINTER(C.c_void_p), C.POINTER(C.c_uint32)]
    lib.sa_rm_seed_and_filter_range.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_uint32, C.c_uint32,
                                                C.POINTER(C.c_void_p), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.sa_host_chunk_seeds.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
    lib.sa_host_chunk_seeds.restype = C.c_size_t
    lib.sa_write_segments.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_int, C.c_uint64, C.c_uint64,
                                      C.POINTER(SaChromTable), C.POINTER(SaChromTable)]
    return lib


def shape_pattern(seed_shape: str) -> str:
    """src/main.cpp:160-178: user seed string -> T/0 pattern."""
    if seed_shape == "12of19":
        return "TTT0T00TT00T0T0TTTT"
    if seed_shape == "14of22":
        return "TTT0T0TT00TT00T0T0TTTT"
    return "".join("T" if c == "1" else "0" for c in seed_shape)


class Backend:
    """One process-wide backend instance (the reference keeps this state in globals)."""

    def __init__(self, lib_path: Path | None = None):
        self.lib = load_library(lib_path)
        self._keepalive = {}
        self.seed_span = 0
        self.seed_weight = 0

    def _check(self, rc: int) -> int:
        if rc < 0:
            raise BackendError(rc, self.lib.sa_last_error().decode(errors="replace"))
        return rc

    # --- reference boundary -------------------------------------------------------------
    def InitializeInterface(self, num_gpu: int = -1, first_device: int = 0) -> int:
        return self._check(self.lib.sa_initialize_interface_at(first_device, num_gpu))

    def GenerateShapePos(self, seed_shape: str) -> int:
        pat = shape_pattern(seed_shape)
        self.seed_span = len(pat)
        self.seed_weight = self._check(self.lib.sa_set_seed_shape(pat.encode()))
        return self.seed_weight

    def InitializeProcessor(self, transition: bool, wga_chunk: int, seed_size: int, sub_mat,
                            xdrop: int, hspthresh: int, noentropy: bool) -> None:
        m = np.ascontiguousarray(sub_mat, dtype=np.int32).reshape(64)
        self._check(self.lib.sa_initialize_processor(int(transition), wga_chunk, seed_size,
                                                     m.ctypes.data_as(C.POINTER(C.c_int)), xdrop,
                                                     hspthresh, int(noentropy)))

    def SendRefWriteRequest(self, seq: np.ndarray, start_addr: int, length: int) -> None:
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        self._keepalive["ref"] = seq
        self._check(self.lib.sa_send_ref(seq.ctypes.data, start_addr, length))

    def GenerateSeedPosTable(self, ref: np.ndarray, start_addr: int, ref_length: int, step: int = 1) -> None:
        ref = np.ascontiguousarray(ref, dtype=np.uint8)
        self._check(self.lib.sa_generate_seed_pos_table(ref.ctypes.data, start_addr, ref_length, step,
                                                        self.seed_span, self.seed_weight))

    def ClearRef(self) -> None:
        self._check(self.lib.sa_clear_ref())

    def SendQueryWriteRequest(self, query: np.ndarray, start_addr: int, length: int, buffer: int) -> None:
        query = np.ascontiguousarray(query, dtype=np.uint8)
        self._check(self.lib.sa_send_query(query.ctypes.data, start_addr, length, buffer))

    def ClearQuery(self, buffer: int) -> None:
        self._check(self.lib.sa_clear_query(buffer))

    def _take(self, out, n) -> np.ndarray:
        cnt = n.value
        arr = np.empty(cnt, dtype=SEGMENT_DTYPE)
        C.memmove(arr.ctypes.data, out.value, cnt * 16)
        self.lib.sa_release_result(out)
        return arr

    def SeedAndFilter(self, seed_offset_vector: np.ndarray, rev: bool, buffer: int) -> np.ndarray:
        """Returns the reference's vector: element 0 = header {0,0,len=#HSPs,score=#hits}."""
        seeds = np.ascontiguousarray(seed_offset_vector, dtype=np.uint64)
        out, n = C.c_void_p(), C.c_uint32()
        self._check(self.lib.sa_seed_and_filter(seeds.ctypes.data, seeds.size, int(rev), buffer,
                                                C.byref(out), C.byref(n)))
        return self._take(out, n)

    def SeedAndFilterRange(self, q_start: int, q_end: int, transition: bool, rev: bool, buffer: int):
        out, n, ns = C.c_void_p(), C.c_uint32(), C.c_uint32()
        self._check(self.lib.sa_seed_and_filter_range(q_start, q_end, int(transition), int(rev), buffer,
                                                      C.byref(out), C.byref(n), C.byref(ns)))
        return self._take(out, n), ns.value

    def host_chunk_seeds(self, seq: np.ndarray, j0: int, j1: int, transition: bool,
                         out: np.ndarray | None = None) -> np.ndarray:
        """src/seeder.cpp:57-74 natively; `out` (uint64, >= (j1-j0)*(1+weight)) may be pinned."""
        if out is None:
            out = np.empty(max(1, (j1 - j0) * (1 + self.seed_weight)), dtype=np.uint64)
        n = self.lib.sa_host_chunk_seeds(seq.ctypes.data, 0, j0, j1, int(transition), out.ctypes.data)
        return out[:n]

    def SeedAndFilterPtr(self, seeds_ptr: int, num_seeds: int, rev: bool, buffer: int) -> np.ndarray:
        out, n = C.c_void_p(), C.c_uint32()
        self._check(self.lib.sa_seed_and_filter(seeds_ptr, num_seeds, int(rev), buffer,
                                                C.byref(out), C.byref(n)))
        return self._take(out, n)

    def write_segments(self, path, hsps: np.ndarray, minus: bool, r_block_start: int, q_block_start: int,
                       ref_chroms: "SaChromTable", query_chroms: "SaChromTable") -> None:
        """src/segment_printer.cpp:72-94 / :125-149 (host only; works without a GPU)."""
        hsps = np.ascontiguousarray(hsps, dtype=SEGMENT_DTYPE)
        self._check(self.lib.sa_write_segments(str(path).encode(), hsps.ctypes.data, hsps.size, int(minus),
                                               r_block_start, q_block_start, C.byref(ref_chroms),
                                               C.byref(query_chroms)))

    # --- repeat-masker variant (repeat_masker_src/seed_filter.h:5-14) ---------------------
    def RmSendQueryWriteRequest(self) -> None:
        self._check(self.lib.sa_rm_send_query())

    def RmClearQuery(self) -> None:
        self._check(self.lib.sa_rm_clear_query())

    def RmSeedAndFilter(self, seed_offset_vector: np.ndarray, rev: bool, ref_start: int, ref_end: int) -> np.ndarray:
        seeds = np.ascontiguousarray(seed_offset_vector, dtype=np.uint64)
        out, n = C.c_void_p(), C.c_uint32()
        self._check(self.lib.sa_rm_seed_and_filter(seeds.ctypes.data, seeds.size, int(rev), ref_start, ref_end,
                                                   C.byref(out), C.byref(n)))
        return self._take(out, n)

    def RmSeedAndFilterRange(self, q_start: int, q_end: int, transition: bool, rev: bool, ref_start: int, ref_end: int):
        out, n, ns = C.c_void_p(), C.c_uint32(), C.c_uint32()
        self._check(self.lib.sa_rm_seed_and_filter_range(q_start, q_end, int(transition), int(rev), ref_start, ref_end,
                                                         C.byref(out), C.byref(n), C.byref(ns)))
        return self._take(out, n), ns.value

    def ShutdownProcessor(self) -> None:
        self._check(self.lib.sa_shutdown_processor())

    # --- knobs / introspection ----------------------------------------------------------
    def set_max_hits(self, n: int) -> None:
        self._check(self.lib.sa_set_max_hits(n))

    def get_max_hits(self) -> int:
        return self.lib.sa_get_max_hits()

    def build_matrix(self, ambiguous: str, xdrop: int) -> np.ndarray:
        m = np.zeros(64, dtype=np.int32)
        self._check(self.lib.sa_build_matrix(ambiguous.encode(), int(xdrop), m.ctypes.data_as(C.POINTER(C.c_int))))
        return m

    def pipeline_run(self, ref_fasta, query_fasta, out_dir, plan_only: bool = False, **kw) -> dict:
        """sa_pipeline_run: the Boost-free whole-genome driver; plan_only = its host-only
        part (sa_pipeline_plan: blocks, intervals, name files; no GPU)."""
        cfg = SaPipelineConfig()
        cfg.ref_fasta, cfg.query_fasta, cfg.out_dir = str(ref_fasta).encode(), str(query_fasta).encode(), str(out_dir).encode()
        keep = []
        for k, v in kw.items():
            if k == "sub_mat":
                arr = np.ascontiguousarray(v, dtype=np.int32)
                keep.append(arr)
                cfg.sub_mat = arr.ctypes.data_as(C.POINTER(C.c_int))
            elif isinstance(v, str):
                setattr(cfg, k, v.encode())
            else:
                setattr(cfg, k, int(v))
        rep = SaPipelineReport()
        fn = self.lib.sa_pipeline_plan if plan_only else self.lib.sa_pipeline_run
        self._check(fn(C.byref(cfg), C.byref(rep)))
        return rep.as_dict()

    def set_filter_kernel(self, k: int) -> int:
        return self.lib.sa_set_filter_kernel(int(k))

    def get_table(self):
        isz, npos = C.c_uint32(), C.c_uint32()
        self._check(self.lib.sa_debug_get_table(C.byref(isz), C.byref(npos), None, None))
        index = np.empty(isz.value, dtype=np.uint32)
        pos = np.empty(max(npos.value, 1), dtype=np.uint32)
        self._check(self.lib.sa_debug_get_table(C.byref(isz), C.byref(npos), index.ctypes.data, pos.ctypes.data))
        return index, pos[: npos.value]

    def get_encoded(self, which: int, buffer: int, length: int) -> np.ndarray:
        out = np.empty(length, dtype=np.uint8)
        self._check(self.lib.sa_debug_get_encoded(which, buffer, out.ctypes.data, length))
        return out

    def stats(self) -> dict:
        s = SaStats()
        self._check(self.lib.sa_get_stats(C.byref(s)))
        return s.as_dict()

    def gpu_calls(self) -> list:
        """SeedAndFilter calls served by each GPU of the pool since reset_stats()."""
        buf = (C.c_uint64 * 64)()
        n = self._check(self.lib.sa_get_gpu_calls(buf, 64))
        return [int(buf[i]) for i in range(min(n, 64))]

    def reset_stats(self) -> None:
        self._check(self.lib.sa_reset_stats())

    def set_profiling(self, on: bool) -> None:
        self._check(self.lib.sa_set_profiling(int(on)))
