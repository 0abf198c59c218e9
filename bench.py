#!/usr/bin/env python
"""bench.py -- Gbp of query per second through seed + filter + ungapped extend (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # the CUDA backend (this repo)
    python bench.py --impl reference --gpus N ...            # LASTZ 1.04.17 (the reference's CPU
                                                             #  seeding path) on the host cores

A *step* is one pass of the hot path over one query block: every SeedAndFilter unit (250 kb
chunk x strand, src/seeder.cpp:48-120) of the block against the resident reference block.
Workload at N=1: BASELINE.json configs[1] restated synthetically: ref = 7
records totalling 100 286 401 bp, query = per-record mutate(0.25) + 5 inversions of 1 Mb, 15 %
soft-masked, defaults (12of19, transitions, xdrop 910, hspthresh 3000, entropy, both strands).
N>1: one process per GPU (torchrun); every rank holds the same reference block + seed position
table and its OWN query block of the same size (weak scaling); no data-path collective exists
 -- NCCL carries only the barrier and the max/sum reductions of the report.

Numbers on the JSON line:
  value     whole-job Gbp/s, query block + table resident in HBM, seed words generated on the
            device (sa_seed_and_filter_range), HSPs copied back (they are the result).
  e2e       the same with HOST buffers through the public C ABI as the library's own driver uses it:
            per step the ASCII query block is uploaded from pinned host memory (sa_send_query),
            every unit is one sa_seed_and_filter_range call, HSPs come back to host memory.
            e2e.vector_abi: the unmodified reference seeder's path -- every unit's seed vector is
            built on the host (sa_host_chunk_seeds == src/seeder.cpp:57-74) and handed to
            sa_seed_and_filter (== g_SeedAndFilter); that leg is bound by the host loop that
            writes 104 bytes per query position.
  roofline  k_filter_hits3 (the dominant kernel: seeding + lookup + bucket expansion + the score
            filter over ALL hits): algorithmic bytes 16*S + 4*H + 64*H + E (SURVEY 8d B_L + B_X)
            per launch / CUDA-event duration of that kernel on its own stream in a serialized pass
            right after the timed region, vs MEASURED_PEAKS hbm.
  cpu_baseline  LASTZ (oracle/_ref/lastz, built from the reference's submodule) on a bounded
            sample of the same workload, one process per host core.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
import zlib
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from segalign_b200 import genome  # noqa: E402

SEED_SHAPE = "12of19"
XDROP, HSPTHRESH = 910, 3000
CE11_RECORDS = [15_072_434, 15_279_421, 13_783_801, 17_493_829, 20_924_180, 17_718_942, 13_794]  # 100 286 401


# ------------------------------------------------------------------------------ workload
def scaled_records(total_mb: float | None):
    if not total_mb:
        return CE11_RECORDS
    f = total_mb * 1e6 / sum(CE11_RECORDS)
    return [max(2000, int(r * f)) for r in CE11_RECORDS]


def make_ref(records, seed=20261017):
    rng = np.random.default_rng(seed)
    chroms = [genome.soft_mask(genome.random_genome(n, rng), 0.15, rng) for n in records]
    return chroms


def make_query(ref_chroms, rank=0, d=0.25, inversions=5, inv_len=1_000_000, seed=20261017):
    rng = np.random.default_rng([seed, 1000 + rank])
    out = []
    for c in ref_chroms:
        q = genome.mutate(c, d, rng)                 # keeps the ref's soft-mask (case is preserved)
        q = genome.soft_mask(np.where(q >= 97, q - 32, q).astype(np.uint8), 0.15, rng)  # own mask
        out.append(q)
    total = sum(c.size for c in out)
    for _ in range(inversions):
        c = out[int(rng.integers(0, len(out)))]
        L = min(inv_len, c.size // 4)
        if L < 100:
            continue
        s = int(rng.integers(0, c.size - L))
        c[s:s + L] = genome.revcomp_ascii(c[s:s + L])
    assert total == sum(c.size for c in out)
    return out


def default_matrix():
    """src/main.cpp:187-268, default (no --ambiguous): SURVEY App. C."""
    m = np.full((8, 8), -1000, dtype=np.int32)
    m[:4, :4] = [[91, -114, -31, -123], [-114, 100, -125, -31], [-31, -125, 100, -114], [-123, -31, -114, 91]]
    m[6, :4] = m[:4, 6] = -100
    m[6, 6] = -100
    m[7, :] = m[:, 7] = -10 * XDROP
    return m.reshape(64)


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            self.path = tempfile.NamedTemporaryFile(prefix="clocks_", suffix=".csv", delete=False).name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.proc.wait()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in Path(self.path).read_text().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------ our arm
def nthreads_for_env(args, world):
    return args.host_threads or max(4, min(16, (os.cpu_count() or 4) // max(1, world)))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from segalign_b200.backend import Backend, shape_pattern

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: anything libraries print there (NCCL's version banner)
    # goes to stderr until the line is written
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the backend has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    records = scaled_records(args.ref_mb)
    ref_chroms = make_ref(records)
    ref = genome.make_blocks(ref_chroms)[0]
    query = genome.make_blocks(make_query(ref_chroms, rank))[0]
    if args.query_mb:
        query = query[: int(args.query_mb * 1e6)]
    pattern = shape_pattern(SEED_SHAPE)
    span = len(pattern)
    units = genome.chunk_list(query.size, span, "both")
    q_rc_ascii = genome.revcomp_ascii(query)
    query_bases = int(query.size)

    os.environ["SEGALIGN_B200_STREAMS"] = str(nthreads_for_env(args, world))
    be = Backend()
    be.InitializeInterface(1, first_device=local)
    be.GenerateShapePos(SEED_SHAPE)
    be.InitializeProcessor(True, genome.DEFAULT_WGA_CHUNK, span, default_matrix(), XDROP, HSPTHRESH, False)
    t0 = time.perf_counter()
    be.SendRefWriteRequest(ref, 0, ref.size)
    t1 = time.perf_counter()
    be.GenerateSeedPosTable(ref, 0, ref.size, 1)
    t2 = time.perf_counter()
    be.SendQueryWriteRequest(query, 0, query.size, 0)
    t3 = time.perf_counter()

    # host callers (the reference's TBB seeder workers): enough to keep PCIe, the host seeding
    # loop and the GPU busy at once; one backend workspace (stream) per caller
    nthreads = args.host_threads or max(4, min(16, (os.cpu_count() or 4) // max(1, world)))
    pool = ThreadPoolExecutor(max_workers=nthreads)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # checksum of checksums over the HSP records of one step (order-independent over the units): the
    # three legs must return the same bytes at full size
    step_crc = {}

    def unit_crc(u, res):
        return (zlib.crc32(res[1:].tobytes()) * (2 * u + 1)) & 0xFFFFFFFFFFFF

    # ---- leg 1: resident inputs, device seeding -------------------------------------------
    def step_resident():
        crcs = [0] * len(units)

        def work(u):
            rev, j0, j1 = units[u]
            res, ns = be.SeedAndFilterRange(j0, j1, True, bool(rev), 0)
            crcs[u] = unit_crc(u, res)
            return res.size - 1
        n = sum(pool.map(work, range(len(units))))
        step_crc["resident"] = sum(crcs) & 0xFFFFFFFFFFFFFFFF
        return n

    # ---- leg 2: reference ABI with host buffers ------------------------------------------
    max_words = genome.DEFAULT_WGA_CHUNK * 13
    pinned = [torch.empty(max_words, dtype=torch.int64, pin_memory=True) for _ in range(nthreads)]
    pinned_np = [p.numpy().view(np.uint64) for p in pinned]
    free_bufs = list(range(nthreads))
    buf_lock = threading.Lock()
    e2e_bytes = {"h2d_handed_over": 0, "d2h": 0}

    def step_e2e():
        be.ClearQuery(0)
        be.SendQueryWriteRequest(query, 0, query.size, 0)   # pageable host ASCII -> HBM, as main.cpp:661
        h2d, d2h = [query.size], [0]
        crcs = [0] * len(units)

        def work(u):
            rev, j0, j1 = units[u]
            with buf_lock:
                b = free_bufs.pop()
            try:
                seeds = be.host_chunk_seeds(q_rc_ascii if rev else query, j0, j1, True, pinned_np[b])
                if seeds.size == 0:
                    return 0
                res = be.SeedAndFilterPtr(seeds.ctypes.data, seeds.size, bool(rev), 0)
            finally:
                with buf_lock:
                    free_bufs.append(b)
            crcs[u] = unit_crc(u, res)
            with buf_lock:
                h2d[0] += seeds.size * 8
                d2h[0] += res.size * 16
            return res.size - 1
        n = sum(pool.map(work, range(len(units))))
        e2e_bytes["h2d_handed_over"], e2e_bytes["d2h"] = h2d[0], d2h[0]
        step_crc["vector_abi"] = sum(crcs) & 0xFFFFFFFFFFFFFFFF
        return n

    # ---- leg 3: the library's own driver path with host buffers (what sa_pipeline_run does per query
    # block): the ASCII block is uploaded from pinned host memory every step, seed words are generated
    # on the device, HSPs come back to host memory.
    query_pinned = torch.empty(query.size, dtype=torch.uint8, pin_memory=True)
    query_pinned.numpy()[:] = query
    query_pinned_np = query_pinned.numpy()
    range_d2h = [0]

    def step_e2e_range():
        be.ClearQuery(0)
        be.SendQueryWriteRequest(query_pinned_np, 0, query.size, 0)
        d2h = [0]
        crcs = [0] * len(units)

        def work(u):
            rev, j0, j1 = units[u]
            res, ns = be.SeedAndFilterRange(j0, j1, True, bool(rev), 0)
            crcs[u] = unit_crc(u, res)
            with buf_lock:
                d2h[0] += res.size * 16
            return res.size - 1
        n = sum(pool.map(work, range(len(units))))
        range_d2h[0] = d2h[0]
        step_crc["e2e"] = sum(crcs) & 0xFFFFFFFFFFFFFFFF
        return n

    def timed(step_fn, steps, warmup):
        for _ in range(warmup):
            step_fn()
            flush.zero_()
        barrier()
        be.reset_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        w0 = time.perf_counter()
        hsps = 0
        for _ in range(steps):
            hsps += step_fn()
            flush.zero_()          # L2 flush between steps (256 MB > 126 MB L2)
        barrier()
        e1.record()
        e1.synchronize()
        wall = time.perf_counter() - w0
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), hsps, be.stats()

    sampler = ClockSampler(local)
    sampler.start()
    ms_res, wall_res, hsps_res, st_res = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop()
    ms_e2e, wall_e2e, hsps_e2e, st_e2e = timed(step_e2e, args.steps, max(1, args.warmup // 3))
    assert hsps_e2e == hsps_res, f"e2e path returned {hsps_e2e} HSPs, resident path {hsps_res}"
    ms_e2r, wall_e2r, hsps_e2r, st_e2r = timed(step_e2e_range, args.steps, max(1, args.warmup // 3))
    assert hsps_e2r == hsps_res, f"range e2e path returned {hsps_e2r} HSPs, resident path {hsps_res}"
    assert step_crc["resident"] == step_crc["vector_abi"] == step_crc["e2e"], f"legs returned different HSP bytes: {step_crc}"

    total_bases = torch.tensor([float(query_bases)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_bases, op=dist.ReduceOp.SUM)
    total_bases = float(total_bases[0])
    value = total_bases * args.steps / (ms_res * 1e-3) / 1e9
    e2e_value = total_bases * args.steps / (ms_e2e * 1e-3) / 1e9
    e2r_value = total_bases * args.steps / (ms_e2r * 1e-3) / 1e9

    # Roofline of the dominant kernel.  Inside the timed region `nthreads` calls are in flight at once,
    # so a per-stream CUDA-event interval there also contains the time the kernel waited for SMs;
    # the launch duration is therefore taken from a serialized pass (one call at a time, same
    # inputs, same kernels) run right after the timed region, with the library's CUDA events
    # recorded on the kernel's own stream.  The kernel's share of the concurrent step is reported too.
    n_probe = min(len(units), args.roofline_launches)
    be.reset_stats()
    torch.cuda.synchronize()
    for u in range(n_probe):
        rev, j0, j1 = units[(u * 7) % len(units)]
        be.SeedAndFilterRange(j0, j1, True, bool(rev), 0)
    torch.cuda.synchronize()
    st_probe = be.stats()
    # E of the byte formula (cells the reference's 32-cell tiles scan beyond the first tile of a
    # direction) is a property of the workload, not of the kernel: the default kernel decides most
    # hits by popcounts and counts only the hits it tile-walks, so E is taken from an untimed pass
    # of the tile-walk-only kernel over the same units.
    prev_kernel = be.set_filter_kernel(2)
    be.reset_stats()
    for u in range(n_probe):
        rev, j0, j1 = units[(u * 7) % len(units)]
        be.SeedAndFilterRange(j0, j1, True, bool(rev), 0)
    torch.cuda.synchronize()
    st_acct = be.stats()
    be.set_filter_kernel(prev_kernel)
    assert st_acct["hits"] == st_probe["hits"] and st_acct["hsps"] == st_probe["hsps"]
    ext_cells_probe = st_acct["ext_cells"]
    ext_per_hit = ext_cells_probe / max(1, st_acct["hits"])
    # DRAM traffic of the dominant kernel per launch: from the committed `ncu --set full` capture of
    # this same command line (profiles/*_k_filter_hits3_ncu_summary.txt; bench.py never runs under ncu)
    traffic, traffic_src = None, None
    if prev_kernel == 3:
        caps = sorted((ROOT / "profiles").glob("*_k_filter_hits3_ncu_summary.txt"))
        if caps:
            tot, n_cap = 0.0, 0
            for line in caps[-1].read_text().splitlines():
                if line.startswith("dram__bytes_read.sum") or line.startswith("dram__bytes_write.sum"):
                    unit = line.split()[1]
                    vals = [float(x) for x in line[line.index("["):].strip("[]").replace("'", "").split(",")]
                    tot += sum(vals) / len(vals) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
                    n_cap += 1
            if n_cap == 2:
                traffic, traffic_src = round(tot), "profiles/" + caps[-1].name
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak = float(peaks.get("hbm_gbs", 6650.0))
    n_launch = max(1, st_probe["calls"])
    # the dominant kernel does lookup + expansion + extension filter in one launch (fused path):
    # B_L + B_X = (16 S + 4 H) + (64 H + E), SURVEY 8d
    lookup_bytes = 16.0 * st_probe["seeds"] + 4.0 * st_probe["hits"]
    alg_bytes = lookup_bytes + 64.0 * st_probe["hits"] + ext_cells_probe
    t_ext = st_probe["ms_prefilter"] * 1e-3
    achieved = alg_bytes / t_ext / 1e9 if t_ext > 0 else 0.0
    step_bytes = 16.0 * st_res["seeds"] + 68.0 * st_res["hits"] + ext_per_hit * st_res["hits"]
    roofline = {"bound": "hbm", "kernel": "k_filter_hits%s" % ("3" if prev_kernel == 3 else ("2" if prev_kernel == 2 else "")), "achieved": round(achieved, 1), "peak": peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650",
                "unit": "GB/s", "frac": round(achieved / peak, 4),
                "algorithmic_bytes_per_launch": round(alg_bytes / n_launch),
                "avg_launch_ms": round(st_probe["ms_prefilter"] / n_launch, 4), "launches": n_launch,
                "measured": "serialized pass of %d launches after the timed region (CUDA events on the kernel's stream)" % n_launch,
                "traffic": traffic, "traffic_source": traffic_src,
                "bytes_formula": "16*S + 4*H (seed lookup, fused into this kernel) + 64*H + E (extension), per rank; "
                                 "E counted by an untimed pass of the tile-walk-only kernel over the same units",
                "filter_kernel": int(prev_kernel),
                "lookup": {"fused_into": "the same kernel", "algorithmic_bytes_per_launch": round(lookup_bytes / n_launch)},
                "whole_step": {"algorithmic_GBps": round(step_bytes / (ms_res * 1e-3) / 1e9, 1),
                               "frac": round(step_bytes / (ms_res * 1e-3) / 1e9 / peak, 4),
                               "note": "all kernels + host gaps of the timed region, rank 0"},
                "serialized_phase_ms_per_launch": {k: round(st_probe[k] / n_launch, 4) for k in
                                                   ("ms_h2d", "ms_prefilter", "ms_extend", "ms_sort", "ms_d2h")}}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = lastz_baseline(ref_chroms, query, budget_s=args.cpu_budget)

    if rank == 0:
        line = {
            "metric": "Gbp of query processed/sec (seed+filter+ungapped-extend)", "value": round(value, 5),
            "unit": "Gbp/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_res / args.steps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/int32 (f64 entropy factor)", "data": "synthetic",
            "config": {"workload": "ce11-vs-cb4-scale synthetic (BASELINE configs[1]): ref %d bp in %d records, "
                                   "query %d bp per GPU, mutate 0.25, 5x1Mb inversions, 15%% soft-masked" %
                                   (ref.size, len(records), query_bases),
                       "seed": SEED_SHAPE, "transition": True, "xdrop": XDROP, "hspthresh": HSPTHRESH,
                       "strand": "both", "wga_chunk": genome.DEFAULT_WGA_CHUNK, "units_per_step": len(units),
                       "host_threads": nthreads, "l2": "256 MB memset between steps; per-step working set >> L2",
                       "parallelism": f"query blocks x{world}, replicated ref+table, no collective"},
            "e2e": {"value": round(e2r_value, 5), "unit": "Gbp/s", "ms_per_step": round(ms_e2r / args.steps, 3),
                    "h2d_bytes_per_step": int(st_e2r["h2d_bytes"] // args.steps), "d2h_bytes_per_step": int(range_d2h[0]),
                    "api": "sa_send_query (ASCII query block from pinned host memory, every step) + sa_seed_and_filter_range "
                           "per unit (seed words generated on the device) + HSPs copied back to host memory: the path of "
                           "the library's own driver (sa_pipeline_run) and of the 5-line seeder change in INTEGRATION.md",
                    "vector_abi": {"value": round(e2e_value, 5), "unit": "Gbp/s", "ms_per_step": round(ms_e2e / args.steps, 3),
                                   "h2d_bytes_per_step": int(st_e2e["h2d_bytes"] // args.steps),
                                   "d2h_bytes_per_step": int(e2e_bytes["d2h"]),
                                   "host_bytes_handed_over_per_step": int(e2e_bytes["h2d_handed_over"]),
                                   "api": "sa_send_query + sa_host_chunk_seeds + sa_seed_and_filter: the unmodified reference "
                                          "seeder's seed-vector ABI (g_SeedAndFilter).  Bound by the HOST: the seeder writes "
                                          "104 bytes per query position and strand (17.9 GB per step on %d threads); the "
                                          "library recognises the canonical vector, uploads its base words (1/13) and "
                                          "rebuilds the variants on the device" % nthreads}},
            "gpu_launches": int(st_res["launches"]),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "clocks": clocks,
            "counters_per_step": {"seeds": st_res["seeds"] // args.steps, "hits": st_res["hits"] // args.steps,
                                  "filter_survivors": st_res["survivors"] // args.steps,
                                  "anchors_pre_dedupe": st_res["anchors_pre_dedupe"] // args.steps,
                                  "hsps": hsps_res // args.steps,
                                  "tile_walked_after_screen": st_res["walked"] // args.steps,
                                  "ext_cells_beyond_first_tile": int(ext_per_hit * st_res["hits"]) // args.steps},
            "rates": {"seed_words_per_s": round(st_res["seeds"] / (ms_res * 1e-3), 1),
                      "hits_per_s": round(st_res["hits"] / (ms_res * 1e-3), 1),
                      "note": "rank 0, resident leg (SURVEY 8d asks for seeds/s and hits/s beside the headline)"},
            "setup_ms": {"ref_upload_encode": round((t1 - t0) * 1e3, 1), "seed_pos_table_build": round((t2 - t1) * 1e3, 1),
                         "query_upload_encode": round((t3 - t2) * 1e3, 1)},
            "wall_ms_per_step": round(wall_res / args.steps, 3),
            "hsp_bytes_checksum": {"value": "%016x" % step_crc["resident"],
                                   "note": "sum over units of crc32(HSP records) * (2u+1); identical for the resident, "
                                           "e2e and vector-ABI legs (asserted)"},
        }
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    be.ShutdownProcessor()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------ LASTZ (reference CPU path)
LASTZ = ROOT / "oracle" / "_ref" / "lastz"
LASTZ_ARGS = ["--seed=12of19", "--transition", "--step=1", "--xdrop=910", "--hspthresh=3000", "--nogapped",
              "--strand=both", "--format=segments"]


def write_fasta(path: Path, records, prefix):
    with open(path, "wb") as f:
        for i, r in enumerate(records):
            f.write(b">%s%d\n" % (prefix.encode(), i))
            f.write(r.tobytes())
            f.write(b"\n")


def lastz_run(ref_fa: Path, pieces, cores: int) -> float:
    """One single-threaded LASTZ process per query piece, `cores` at a time (the reference's own
    wrapper parallelises LASTZ this way, scripts/run_segalign:115).  Returns wall seconds."""
    t0 = time.perf_counter()
    procs = []
    for p in pieces:
        procs.append(subprocess.Popen([str(LASTZ), f"{ref_fa}[multiple]", str(p), *LASTZ_ARGS],
                                      stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
    for p in procs:
        p.wait()
    return time.perf_counter() - t0


def lastz_baseline(ref_chroms, query, budget_s=20.0, cores=None):
    """Bounded sample: every core aligns its own query piece against the full reference block.
    LASTZ's per-process target loading + seed-position-table build is timed separately with a
    100-base query and subtracted (the metric counts query throughput with the index resident, as
    for the GPU arm whose table build is reported under setup_ms)."""
    if not LASTZ.exists():
        return {"value": None, "unit": "Gbp/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/lastz missing"}
    cores = cores or os.cpu_count() or 1
    tmp = Path(tempfile.mkdtemp(prefix="sa_lastz_"))
    ref_fa = tmp / "ref.fa"
    write_fasta(ref_fa, ref_chroms, "r")
    rng = np.random.default_rng(7)
    tiny = [tmp / f"tiny{i}.fa" for i in range(cores)]
    for t in tiny:
        write_fasta(t, [genome.random_genome(100, rng)], "t")
    t_setup = min(lastz_run(ref_fa, tiny, cores), lastz_run(ref_fa, tiny, cores))
    # calibrate on a small piece, then size the sample for ~budget_s of work per core
    piece_len = 20_000
    starts = rng.integers(0, max(1, query.size - piece_len), size=cores)
    cal = []
    for i, s in enumerate(starts):
        p = tmp / f"cal{i}.fa"
        write_fasta(p, [np.where(query[s:s + piece_len] == ord("&"), ord("N"), query[s:s + piece_len]).astype(np.uint8)], "q")
        cal.append(p)
    t_cal = max(1e-3, lastz_run(ref_fa, cal, cores) - t_setup)
    piece_len = int(min(query.size // cores, max(piece_len, piece_len * budget_s / t_cal)))
    starts = (np.arange(cores) * (query.size // cores)).astype(np.int64)
    pieces = []
    for i, s in enumerate(starts):
        p = tmp / f"piece{i}.fa"
        write_fasta(p, [np.where(query[s:s + piece_len] == ord("&"), ord("N"), query[s:s + piece_len]).astype(np.uint8)], "q")
        pieces.append(p)
    t_run = lastz_run(ref_fa, pieces, cores)
    for f in tmp.glob("*"):
        f.unlink()
    tmp.rmdir()
    subtracted = t_run - t_setup > 0.25 * t_run
    work = t_run - t_setup if subtracted else t_run
    bases = piece_len * cores
    return {"value": round(bases / work / 1e9, 6), "unit": "Gbp/s", "cores": cores, "kind": "reference",
            "sample": "LASTZ 1.04.17 (reference submodule): %d processes x %d bp query pieces vs the full %d bp "
                      "reference; %.1f s wall %s %.1f s per-process target load + table build"
                      % (cores, piece_len, sum(c.size for c in ref_chroms), t_run,
                         "minus" if subtracted else "(not reduced by the)", t_setup),
            "per_core": round(bases / work / 1e9 / cores, 7), "wall_s": round(t_run, 2), "setup_s": round(t_setup, 2)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    records = scaled_records(args.ref_mb)
    ref_chroms = make_ref(records)
    query = genome.make_blocks(make_query(ref_chroms, 0))[0]
    if args.query_mb:
        query = query[: int(args.query_mb * 1e6)]
    cores = os.cpu_count() or 1
    vals, last = [], None
    n_total = args.steps + args.warmup
    per = max(5.0, args.cpu_budget * 3 / max(1, n_total))
    t0 = time.perf_counter()
    for i in range(n_total):
        last = lastz_baseline(ref_chroms, query, budget_s=per, cores=cores)
        if i >= args.warmup:
            vals.append(last["value"])
    wall = time.perf_counter() - t0
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": "Gbp of query processed/sec (seed+filter+ungapped-extend)",
            "value": round(v, 6), "unit": "Gbp/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(wall / n_total * 1e3, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": "ce11-vs-cb4-scale synthetic (BASELINE configs[1]); each step = bounded sample: "
                                   + last["sample"]},
            "cpu_baseline": {**last, "value": round(v, 6)},
            "e2e": {"value": round(v, 6), "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-mb", type=float, default=None, help="scale the reference block (testing only)")
    ap.add_argument("--query-mb", type=float, default=None, help="truncate the query block (testing only)")
    ap.add_argument("--host-threads", type=int, default=0, help="0 = auto: min(16, cores / ranks)")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--roofline-launches", type=int, default=160)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
