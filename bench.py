#!/usr/bin/env python
"""bench.py -- Gbp of query per second through seed + filter + ungapped extend (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # the CUDA backend (this repo)
    python bench.py --impl reference --gpus N ...            # LASTZ 1.04.17 (the reference's CPU
                                                             #  seeding path) on the host cores
    python bench.py --gpus N --inproc                        # ONE process, N GPUs in the library's
                                                             #  own pool (the reference's model)

A *step* is one pass of the hot path over one query block: every SeedAndFilter unit (250 kb
chunk x strand, src/seeder.cpp:48-120) of the block against the resident reference block.

Workload (default): BASELINE.json configs[2] restated synthetically: the
reference block is 5 records x 100 Mb of i.i.d. uniform ACGT (one 500 000 004-base block with its
'&' separators), no masking; the query is mutate(0.40) of the reference.  One step processes a
100 Mb slice of that query (`--query-mb`; the full 500 Mb query is five such slices, and
throughput per query base does not depend on which slice: every query position meets the whole
500 Mb table -- ~775 seed hits per query base over both strands, 5x the ce11-scale workload).
`--workload ce11` selects BASELINE configs[1] (100 286 401 bp in 7 records,
mutate 0.25 + inversions, 15 % soft-masked), the round-1 workload.
N>1 under torchrun: one process per GPU; every rank holds the same reference block + seed
position table and its OWN query slice (weak scaling); no data-path collective exists
-- NCCL carries only the barrier and the max/sum reductions of the report.  `--strong` under torchrun: all
ranks hold the SAME query block and each runs its static share of the block's SeedAndFilter calls
(segalign_b200/sharding.py); rank 0 then runs every call itself and checks that the union of the
shares is byte-identical.

Numbers on the JSON line:
  value     whole-job Gbp/s, query block + table resident in HBM, seed words generated on the
            device (sa_seed_and_filter_range), HSPs copied back (they are the result).
  e2e       the same with HOST buffers through the public C ABI as the library's own driver uses it:
            per step the ASCII query block is uploaded from pinned host memory (sa_send_query),
            every unit is one sa_seed_and_filter_range call, HSPs come back to host memory.
            e2e.vector_abi: the unmodified reference seeder's path -- every unit's seed vector is
            built on the host (sa_host_chunk_seeds == src/seeder.cpp:57-74) and handed to
            sa_seed_and_filter (== g_SeedAndFilter).
  roofline  k_filter_hits3 (the dominant kernel: seeding + lookup + bucket expansion + the score
            filter over ALL hits): algorithmic bytes 16*S + 4*H + 64*H + E (SURVEY 8d B_L + B_X)
            per launch / CUDA-event duration of that kernel on its own stream in a serialized pass
            right after the timed region, vs MEASURED_PEAKS hbm.
  reference_gpu  the reference's OWN kernels (oracle/_ref/oracle_runner = the unmodified
            seed_filter.cu etc. rebuilt for sm_100a) on a slice of the SAME inputs on the same
            GPU: seconds inside SeedAndFilter, and a byte-compare of every call's records with
            this backend's (the bench fails on a mismatch).
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

from segalign_b200 import genome, sharding  # noqa: E402

SEED_SHAPE = "12of19"
XDROP, HSPTHRESH = 910, 3000
CE11_RECORDS = [15_072_434, 15_279_421, 13_783_801, 17_493_829, 20_924_180, 17_718_942, 13_794]  # 100 286 401
SYN500_RECORDS = [100_000_000] * 5
METRIC = "Gbp of query processed/sec (seed+filter+ungapped-extend)"


# ------------------------------------------------------------------------------ workloads
class Workload:
    """ref_chroms (list of ASCII arrays), ref block, and the per-rank query block of one step."""

    def __init__(self, args, rank: int):
        self.name = args.workload
        if self.name == "syn500":
            recs = SYN500_RECORDS
            if args.ref_mb:
                recs = [max(2000, int(args.ref_mb * 1e6 / 5))] * 5
            rng = np.random.default_rng(20261017)
            self.ref_chroms = [genome.random_genome(n, rng) for n in recs]
            self.ref = genome.make_blocks(self.ref_chroms)[0]
            qlen = int((args.query_mb or 100.0) * 1e6)
            # this rank's slice of mutate(0.40)(reference): inside one record, a different place per rank
            rec = self.ref_chroms[rank % len(self.ref_chroms)]
            qlen = min(qlen, rec.size)
            span = rec.size - qlen
            off = ((rank // len(self.ref_chroms)) * 37_000_003) % max(1, span + 1) if span > 0 else 0
            qrng = np.random.default_rng([20261017, 2000 + rank])
            self.query = genome.mutate(rec[off:off + qlen], 0.40, qrng)
            self.label = ("synthetic 500 Mb x 500 Mb, 40 %% divergence (BASELINE configs[2]): ref block %d bp in %d records "
                          "(i.i.d. ACGT, unmasked), query = mutate(0.40) of the reference, %d bp slice per GPU per step"
                          % (self.ref.size, len(recs), self.query.size))
            self.config_id = "configs[2]"
        else:
            recs = CE11_RECORDS
            if args.ref_mb:
                f = args.ref_mb * 1e6 / sum(CE11_RECORDS)
                recs = [max(2000, int(r * f)) for r in CE11_RECORDS]
            self.ref_chroms = make_ce11_ref(recs)
            self.ref = genome.make_blocks(self.ref_chroms)[0]
            self.query = genome.make_blocks(make_ce11_query(self.ref_chroms, rank))[0]
            if args.query_mb:
                self.query = self.query[: int(args.query_mb * 1e6)]
            self.label = ("ce11-vs-cb4-scale synthetic (BASELINE configs[1]): ref %d bp in %d records, query %d bp per GPU, "
                          "mutate 0.25, 5x1Mb inversions, 15 %% soft-masked" % (self.ref.size, len(recs), self.query.size))
            self.config_id = "configs[1]"


def make_ce11_ref(records, seed=20261017):
    rng = np.random.default_rng(seed)
    return [genome.soft_mask(genome.random_genome(n, rng), 0.15, rng) for n in records]


def make_ce11_query(ref_chroms, rank=0, d=0.25, inversions=5, inv_len=1_000_000, seed=20261017):
    rng = np.random.default_rng([seed, 1000 + rank])
    out = []
    for c in ref_chroms:
        q = genome.mutate(c, d, rng)                 # keeps the ref's soft-mask (case is preserved)
        q = genome.soft_mask(np.where(q >= 97, q - 32, q).astype(np.uint8), 0.15, rng)  # own mask
        out.append(q)
    for _ in range(inversions):
        c = out[int(rng.integers(0, len(out)))]
        L = min(inv_len, c.size // 4)
        if L < 100:
            continue
        s = int(rng.integers(0, c.size - L))
        c[s:s + L] = genome.revcomp_ascii(c[s:s + L])
    return out


# kept under their round-1 names: tests/golden/check_large.py builds its worm-like piece from them
def scaled_records(total_mb):
    if not total_mb:
        return CE11_RECORDS
    f = total_mb * 1e6 / sum(CE11_RECORDS)
    return [max(2000, int(r * f)) for r in CE11_RECORDS]


make_ref = make_ce11_ref
make_query = make_ce11_query


def default_matrix():
    """src/main.cpp:187-268, default (no --ambiguous): SURVEY App. C."""
    m = np.full((8, 8), -1000, dtype=np.int32)
    m[:4, :4] = [[91, -114, -31, -123], [-114, 100, -125, -31], [-31, -125, 100, -114], [-123, -31, -114, 91]]
    m[6, :4] = m[:4, 6] = -100
    m[6, 6] = -100
    m[7, :] = m[:, 7] = -10 * XDROP
    return m.reshape(64)


def iupac_matrix():
    """src/main.cpp:200-203,229-250 with --ambiguous=iupac: N and the other IUPAC letters score 0 against
    everything below them and against themselves."""
    m = default_matrix().reshape(8, 8).copy()
    m[5, :5] = m[:5, 5] = 0
    m[5, 5] = 0
    m[6, :6] = m[:6, 6] = 0
    m[6, 6] = 0
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
def host_threads(args, world: int, gpus_in_process: int = 1) -> int:
    if args.host_threads:
        return args.host_threads
    cores = os.cpu_count() or 4
    return max(4, min(16 * gpus_in_process, cores // max(1, world)))


def unit_crc(u, res):
    return (zlib.crc32(res[1:].tobytes()) * (2 * u + 1)) & 0xFFFFFFFFFFFF


def peaks_hbm():
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        return float(json.loads(pk.read_text()).get("hbm_gbs", 6650.0)), "MEASURED_PEAKS.json hbm_gbs (measured)"
    return 6650.0, "fallback 6650 (B200_PROFILING.md)"


def committed_traffic(workload_name: str):
    """DRAM bytes per launch of the dominant kernel from the newest committed `ncu --set full` capture of
    this workload (profiles/*_<workload>_k_filter_hits3_ncu_summary.txt; bench.py never runs under ncu)."""
    caps = sorted((ROOT / "profiles").glob(f"*_{workload_name}_k_filter_hits3_ncu_summary.txt"))
    if not caps and workload_name == "ce11":
        caps = sorted(p for p in (ROOT / "profiles").glob("*_k_filter_hits3_ncu_summary.txt") if "syn500" not in p.name)
    if not caps:
        return None, None
    tot, n_cap = 0.0, 0
    for line in caps[-1].read_text().splitlines():
        if line.startswith("dram__bytes_read.sum") or line.startswith("dram__bytes_write.sum"):
            unit = line.split()[1]
            vals = [float(x) for x in line[line.index("["):].strip("[]").replace("'", "").split(",")]
            tot += sum(vals) / len(vals) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            n_cap += 1
    return (round(tot), "profiles/" + caps[-1].name) if n_cap == 2 else (None, None)


def setup_backend(be, args, wl, nthreads_per_gpu, n_gpus_in_process=1, first_device=0):
    from segalign_b200.backend import shape_pattern
    os.environ["SEGALIGN_B200_STREAMS"] = str(nthreads_per_gpu)
    be.InitializeInterface(n_gpus_in_process, first_device=first_device)
    be.GenerateShapePos(SEED_SHAPE)
    span = len(shape_pattern(SEED_SHAPE))
    be.InitializeProcessor(True, genome.DEFAULT_WGA_CHUNK, span, default_matrix(), XDROP, HSPTHRESH, False)
    t0 = time.perf_counter()
    be.SendRefWriteRequest(wl.ref, 0, wl.ref.size)
    t1 = time.perf_counter()
    be.GenerateSeedPosTable(wl.ref, 0, wl.ref.size, 1)
    t2 = time.perf_counter()
    be.SendQueryWriteRequest(wl.query, 0, wl.query.size, 0)
    t3 = time.perf_counter()
    return span, {"ref_upload_encode": round((t1 - t0) * 1e3, 1), "seed_pos_table_build": round((t2 - t1) * 1e3, 1),
                  "query_upload_encode": round((t3 - t2) * 1e3, 1)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from segalign_b200.backend import Backend

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
    inproc = args.inproc and world == 1
    n_local = args.gpus if inproc else 1
    if inproc and torch.cuda.device_count() < n_local:
        raise SystemExit(f"bench.py --inproc: {n_local} GPUs requested, {torch.cuda.device_count()} visible")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    strong = bool(args.strong) and world > 1
    wl = Workload(args, 0 if strong else rank)
    ref, query = wl.ref, wl.query
    nthreads = host_threads(args, world, n_local)
    be = Backend()
    span, setup_ms = setup_backend(be, args, wl, max(2, (nthreads + n_local - 1) // n_local), n_local, first_device=local)
    all_units = genome.chunk_list(query.size, span, "both")
    # --strong under torchrun: every rank holds the SAME query block and owns a static, contiguous share of
    # its SeedAndFilter calls (segalign_b200/sharding.py); the calls are independent, so the union of the
    # ranks' outputs is the single-GPU output (checked below against rank 0 running every call).
    unit_ids = list(sharding.shard_units(len(all_units), rank, world)) if strong else list(range(len(all_units)))
    units = [all_units[i] for i in unit_ids]
    q_rc_ascii = genome.revcomp_ascii(query)
    query_bases = int(query.size)

    # host callers (the reference's TBB seeder workers): enough to keep PCIe, the host seeding
    # loop and the GPU busy at once; one backend workspace (stream) per caller
    pool = ThreadPoolExecutor(max_workers=nthreads)
    flush = [torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local + i}") for i in range(n_local)]

    def flush_l2():
        for f in flush:
            f.zero_()          # L2 flush between steps (256 MB > 126 MB L2)

    def sync_all():
        for i in range(n_local):
            torch.cuda.synchronize(local + i)

    def barrier():
        sync_all()
        if world > 1:
            dist.barrier()
        sync_all()

    # checksum of checksums over the HSP records of one step (order-independent over the units): the
    # three legs must return the same bytes at full size
    step_crc = {}

    # ---- leg 1: resident inputs, device seeding -------------------------------------------
    def step_resident():
        crcs = [0] * len(units)

        def work(u):
            rev, j0, j1 = units[u]
            res, ns = be.SeedAndFilterRange(j0, j1, True, bool(rev), 0)
            crcs[u] = unit_crc(unit_ids[u], res)
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

    def step_vector_abi():
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
            crcs[u] = unit_crc(unit_ids[u], res)
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
            crcs[u] = unit_crc(unit_ids[u], res)
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
            flush_l2()
        barrier()
        be.reset_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        w0 = time.perf_counter()
        hsps = 0
        for _ in range(steps):
            hsps += step_fn()
            flush_l2()
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
    gpu_calls_res = be.gpu_calls()
    ms_e2r, wall_e2r, hsps_e2r, st_e2r = timed(step_e2e_range, args.steps, max(1, args.warmup // 3))
    assert hsps_e2r == hsps_res, f"range e2e path returned {hsps_e2r} HSPs, resident path {hsps_res}"
    vsteps = max(1, min(args.steps, args.vector_steps))
    ms_vec, wall_vec, hsps_vec, st_vec = timed(step_vector_abi, vsteps, 1)
    assert hsps_vec * args.steps == hsps_res * vsteps, f"vector-ABI path returned {hsps_vec} HSPs in {vsteps} steps, resident path {hsps_res} in {args.steps}"
    assert step_crc["resident"] == step_crc["vector_abi"] == step_crc["e2e"], f"legs returned different HSP bytes: {step_crc}"

    total_bases = torch.tensor([float(query_bases)], dtype=torch.float64, device="cuda")
    if world > 1 and not strong:
        dist.all_reduce(total_bases, op=dist.ReduceOp.SUM)
    total_bases = float(total_bases[0])
    strong_check = None
    if strong:
        # union of the ranks' shares == one GPU running every call (same bytes, unit by unit)
        parts = [None] * world
        dist.all_gather_object(parts, (step_crc["resident"], hsps_res // args.steps))
        if rank == 0:
            crc_all, hsps_all = 0, 0
            for u, (rev, j0, j1) in enumerate(all_units):
                res, _ = be.SeedAndFilterRange(j0, j1, True, bool(rev), 0)
                crc_all += unit_crc(u, res)
                hsps_all += res.size - 1
            crc_all &= 0xFFFFFFFFFFFFFFFF
            crc_sum = sum(c for c, _ in parts) & 0xFFFFFFFFFFFFFFFF
            assert crc_sum == crc_all and sum(h for _, h in parts) == hsps_all, \
                f"sharded ranks returned different HSP bytes than one GPU: {crc_sum:#x} vs {crc_all:#x}"
            strong_check = {"identical_to_one_gpu": True, "units": len(all_units), "hsps": hsps_all}
        dist.barrier()
    value = total_bases * args.steps / (ms_res * 1e-3) / 1e9
    e2r_value = total_bases * args.steps / (ms_e2r * 1e-3) / 1e9
    vec_value = total_bases * vsteps / (ms_vec * 1e-3) / 1e9

    # Roofline of the dominant kernel.  Inside the timed region `nthreads` calls are in flight at once,
    # so a per-stream CUDA-event interval there also contains the time the kernel waited for SMs;
    # the launch duration is therefore taken from a serialized pass (one call at a time, same
    # inputs, same kernels) run right after the timed region, with the library's CUDA events
    # recorded on the kernel's own stream.  The kernel's share of the concurrent step is reported too.
    n_probe = min(len(units), args.roofline_launches)
    be.set_profiling(True)
    be.reset_stats()
    sync_all()
    for u in range(n_probe):
        rev, j0, j1 = units[(u * 7) % len(units)]
        be.SeedAndFilterRange(j0, j1, True, bool(rev), 0)
    sync_all()
    st_probe = be.stats()
    # E of the byte formula (cells the reference's 32-cell tiles scan beyond the first tile of a
    # direction) is a property of the workload, not of the kernel: the default kernel decides most
    # hits by popcounts and counts only the hits it tile-walks, so E is taken from an untimed pass
    # of the tile-walk-only kernel over (a part of) the same units.
    n_acct = min(n_probe, args.acct_launches)
    prev_kernel = be.set_filter_kernel(2)
    be.reset_stats()
    for u in range(n_acct):
        rev, j0, j1 = units[(u * 7) % len(units)]
        be.SeedAndFilterRange(j0, j1, True, bool(rev), 0)
    sync_all()
    st_acct = be.stats()
    be.set_filter_kernel(prev_kernel)
    be.set_profiling(False)
    ext_per_hit = st_acct["ext_cells"] / max(1, st_acct["hits"])
    traffic, traffic_src = committed_traffic(wl.name) if prev_kernel == 3 else (None, None)
    peak, peak_src = peaks_hbm()
    n_launch = max(1, st_probe["calls"])
    # the dominant kernel does lookup + expansion + extension filter in one launch (fused path):
    # B_L + B_X = (16 S + 4 H) + (64 H + E), SURVEY 8d
    lookup_bytes = 16.0 * st_probe["seeds"] + 4.0 * st_probe["hits"]
    alg_bytes = lookup_bytes + 64.0 * st_probe["hits"] + ext_per_hit * st_probe["hits"]
    t_ext = st_probe["ms_prefilter"] * 1e-3
    achieved = alg_bytes / t_ext / 1e9 if t_ext > 0 else 0.0
    step_bytes = 16.0 * st_res["seeds"] + 68.0 * st_res["hits"] + ext_per_hit * st_res["hits"]
    roofline = {"bound": "hbm", "kernel": "k_filter_hits3" if prev_kernel == 3 else "k_filter_hits2",
                "achieved": round(achieved, 1), "peak": peak, "peak_source": peak_src,
                "unit": "GB/s", "frac": round(achieved / peak, 4),
                "algorithmic_bytes_per_launch": round(alg_bytes / n_launch),
                "avg_launch_ms": round(st_probe["ms_prefilter"] / n_launch, 4), "launches": n_launch,
                "measured": "serialized pass of %d launches after the timed region (CUDA events on the kernel's stream)" % n_launch,
                "traffic": traffic, "traffic_source": traffic_src,
                "bytes_formula": "16*S + 4*H (seed lookup, fused into this kernel) + 64*H + E (extension), per rank; "
                                 "E/H = %.2f counted by an untimed pass of the tile-walk-only kernel over %d of the same units"
                                 % (ext_per_hit, n_acct),
                "filter_kernel": int(prev_kernel),
                "lookup": {"fused_into": "the same kernel", "algorithmic_bytes_per_launch": round(lookup_bytes / n_launch)},
                "whole_step": {"algorithmic_GBps": round(step_bytes / (ms_res * 1e-3) / 1e9, 1),
                               "frac": round(step_bytes / (ms_res * 1e-3) / 1e9 / peak, 4),
                               "note": "all kernels + host gaps of the timed region, rank 0"},
                "serialized_phase_ms_per_launch": {k: round(st_probe[k] / n_launch, 4) for k in
                                                   ("ms_h2d", "ms_prefilter", "ms_extend", "ms_sort", "ms_d2h")}}

    reference_gpu = None
    if rank == 0 and world == 1 and not inproc and not args.no_reference_gpu:
        reference_gpu = reference_gpu_leg(be, wl, span, args)

    extra = None
    if rank == 0 and world == 1 and not inproc and not args.no_extra:
        extra = extra_workloads(be, args, pool, nthreads)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            be.ShutdownProcessor()   # LASTZ wants the host cores; nothing of ours runs beside it
        except Exception:  # noqa: BLE001 -- already down (a secondary workload gave up half-way)
            pass
        cpu_baseline = lastz_baseline(wl, budget_s=args.cpu_budget)

    if rank == 0:
        par = (f"one process, {n_local} GPUs in the library's pool (replicated ref+table, chunks handed to whichever GPU has a free stream)"
               if inproc else
               f"ONE query block, its SeedAndFilter calls split statically over {world} ranks (sharding.shard_units), replicated ref+table, no collective"
               if strong else f"query blocks x{world}, replicated ref+table, no collective")
        line = {
            "metric": METRIC, "value": round(value, 5),
            "unit": "Gbp/s", "n_gpus": world * n_local, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_res / args.steps, 3), "higher_is_better": True,
            "scaling": "strong" if (inproc or strong) else "weak",
            "vs_baseline": None, "dtype": "u8/int32 (f64 entropy factor)", "data": "synthetic",
            "config": {"workload": wl.label, "baseline_config": wl.config_id,
                       "seed": SEED_SHAPE, "transition": True, "xdrop": XDROP, "hspthresh": HSPTHRESH,
                       "strand": "both", "wga_chunk": genome.DEFAULT_WGA_CHUNK, "units_per_step": len(units),
                       "host_threads": nthreads, "l2": "256 MB memset between steps; per-step working set >> L2",
                       "parallelism": par},
            "e2e": {"value": round(e2r_value, 5), "unit": "Gbp/s", "ms_per_step": round(ms_e2r / args.steps, 3),
                    "h2d_bytes_per_step": int(st_e2r["h2d_bytes"] // args.steps), "d2h_bytes_per_step": int(range_d2h[0]),
                    "api": "sa_send_query (ASCII query block from pinned host memory, every step) + sa_seed_and_filter_range "
                           "per unit (seed words generated on the device) + HSPs copied back to host memory: the path of "
                           "the library's own driver (sa_pipeline_run) and of the 5-line seeder change in INTEGRATION.md",
                    "vector_abi": {"value": round(vec_value, 5), "unit": "Gbp/s", "ms_per_step": round(ms_vec / vsteps, 3),
                                   "steps": vsteps,
                                   "h2d_bytes_per_step": int(st_vec["h2d_bytes"] // vsteps),
                                   "d2h_bytes_per_step": int(e2e_bytes["d2h"]),
                                   "host_bytes_handed_over_per_step": int(e2e_bytes["h2d_handed_over"]),
                                   "api": "sa_send_query + sa_host_chunk_seeds + sa_seed_and_filter: the unmodified reference "
                                          "seeder's seed-vector ABI (g_SeedAndFilter); the host writes 104 bytes per query "
                                          "position and strand on %d threads, the library recognises the canonical vector, "
                                          "uploads its base words (1/13) and rebuilds the variants on the device" % nthreads}},
            "gpu_launches": int(st_res["launches"]),
            "roofline": roofline,
            "reference_gpu": reference_gpu,
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
            "setup_ms": setup_ms,
            "wall_ms_per_step": round(wall_res / args.steps, 3),
            "hsp_bytes_checksum": {"value": "%016x" % step_crc["resident"],
                                   "note": "sum over units of crc32(HSP records) * (2u+1); identical for the resident, "
                                           "e2e and vector-ABI legs (asserted)"},
            "extra": extra,
        }
        if inproc:
            line["mode"] = "inproc"
            line["calls_per_gpu"] = gpu_calls_res
        if strong:
            line["mode"] = "torchrun-strong"
            line["sharded_vs_one_gpu"] = strong_check
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    try:
        be.ShutdownProcessor()
    except Exception:
        pass
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------ the reference's own GPU kernels
ORACLE_RUNNER = ROOT / "oracle" / "_ref" / "oracle_runner"


def reference_gpu_leg(be, wl, span, args):
    return reference_gpu_compare(be, wl.ref, wl.query, span, args, True, default_matrix(), args.reference_gpu_mb)


def reference_gpu_compare(be, ref, query, span, args, transition, matrix, slice_mb):
    """SURVEY 8d "reference GPU timing" / BASELINE.md B1 on the bench workload: oracle_runner (the
    reference's unmodified kernels, sm_100a build) and this backend run the same SeedAndFilter calls
    -- the full reference block, a `--reference-gpu-mb` slice of the bench query as the query block;
    every call's records are byte-compared, both sides' time inside SeedAndFilter is reported."""
    import struct
    from tests import harness as H   # file formats + record comparison only; nothing of oracle/ is imported
    if not ORACLE_RUNNER.exists():
        return {"unavailable": "oracle/_ref/oracle_runner not built (needs /root/reference at build time)"}
    qn = int(min(query.size, slice_mb * 1e6))
    q = np.ascontiguousarray(query[:qn])
    work = Path(tempfile.mkdtemp(prefix="sa_refgpu_"))
    try:
        cf, of = work / "slice.case", work / "slice.out"
        with open(cf, "wb") as f:   # SACASE01 (oracle/ref_driver.cpp:read_case), the bench's own parameters
            shape = SEED_SHAPE.encode()
            f.write(b"SACASE01" + struct.pack("<I", len(shape)) + shape)
            f.write(struct.pack("<iIiiiIIii", int(bool(transition)), 1, XDROP, HSPTHRESH, 0, genome.DEFAULT_WGA_CHUNK,
                                genome.DEFAULT_LASTZ_INTERVAL, 0, 0))
            f.write(np.asarray(matrix).astype("<i4").tobytes())
            f.write(struct.pack("<Q", ref.size)); f.write(ref.tobytes())
            f.write(struct.pack("<Q", q.size)); f.write(q.tobytes())
        t0 = time.perf_counter()
        p = subprocess.run([str(ORACLE_RUNNER), str(cf), str(of)], stderr=subprocess.PIPE, stdout=subprocess.DEVNULL,
                           timeout=args.reference_gpu_timeout)
        t_total = time.perf_counter() - t0
        if p.returncode != 0:
            return {"unavailable": "oracle_runner exit %d: %s" % (p.returncode, p.stderr.decode(errors="replace")[-300:])}
        dump = H.read_dump(of)
    except subprocess.TimeoutExpired:
        return {"unavailable": "oracle_runner exceeded %d s on a %d bp slice" % (args.reference_gpu_timeout, qn)}
    finally:
        for f in work.glob("*"):
            f.unlink()
        work.rmdir()
    want = []
    for rev, cs, ce, ns, tot, nh, segs in dump.calls:
        res = np.zeros(segs.size + 1, dtype=H.SEGMENT_DTYPE)
        res[0]["len"], res[0]["score"] = tot, np.uint32(nh).view(np.int32)
        res[1:] = segs
        want.append((rev, cs, ce, ns, res))
    # the same calls on this backend: the slice goes into the second query slot as its own block
    be.SendQueryWriteRequest(q, 0, q.size, 1)
    got, t_calls = [], 0.0
    for rev, j0, j1 in genome.chunk_list(q.size, span, "both"):
        t1 = time.perf_counter()
        res, ns = be.SeedAndFilterRange(j0, j1, bool(transition), bool(rev), 1)
        t_calls += time.perf_counter() - t1
        if ns:
            got.append((rev, j0, j1, ns, res))
    be.ClearQuery(1)
    H.assert_calls_equal(got, want, "bench workload slice: backend vs the reference's own kernels")  # fails the bench
    ref_s = float(dump.times[4])
    return {"seconds": round(ref_s, 3), "ours_seconds": round(t_calls, 4), "speedup": round(ref_s / max(1e-9, t_calls), 1),
            "identical": True, "calls": len(want), "hits": int(dump.counters[1]), "hsps": int(dump.counters[2]),
            "query_slice_bp": qn, "ref_bp": int(ref.size),
            "gbp_per_s": round(qn / ref_s / 1e9, 6) if ref_s > 0 else None,
            "reference_table_build_s": round(float(dump.times[1]), 3),
            "reference_host_seedgen_s": round(float(dump.times[3]), 3),
            "reference_seeder_positions_per_s_per_thread": round(2.0 * qn / max(1e-9, float(dump.times[3])), 1),
            "reference_seeder_bound_gbp_per_s": round(host_threads(args, 1) * qn / max(1e-9, float(dump.times[3])) / 1e9, 4),
            "reference_seeder_note": "src/seeder.cpp:57-74 (GetKmerIndexAtPos + push_back per seed word) as the runner executes "
                                     "it, one thread; x %d threads = the most query the UNMODIFIED host loop can hand to any "
                                     "backend through g_SeedAndFilter on this box (compare e2e.vector_abi)" % host_threads(args, 1),
            "runner_wall_s": round(t_total, 1),
            "what": "oracle/_ref/oracle_runner = the reference's unmodified seed_filter.cu / seed_pos_table.cu / "
                    "seed_filter_interface.cu (sm_100a build) on the same GPU, one call at a time through g_SeedAndFilter; "
                    "`seconds` = time inside its SeedAndFilter calls (host seed vectors excluded); ours = the same calls, "
                    "serialized, through sa_seed_and_filter_range; every call's records byte-identical"}


# ------------------------------------------------------------------------------ secondary workloads (not the headline)
def make_chr1_pair(n=248_000_000, query_lo=60_000_000, query_n=100_000_000, seed=20261018):
    """configs[3]-scale synthetic pair: one reference record (half soft-masked, an 18 Mb and forty 50 kb runs of N,
    IUPAC letters at 1e-5) and a slice of its 40 %-diverged copy with its own masking and N runs."""
    rng = np.random.default_rng(seed)
    base = genome.random_genome(n, rng)
    q = genome.mutate(base[query_lo:query_lo + query_n], 0.40, rng)
    ref = genome.soft_mask(base, 0.5, rng)
    ref[n // 2 - n // 28:n // 2 + n // 28] = ord("N")
    ref = genome.insert_runs(ref, b"N", 40, 50_000, rng)
    ref = genome.sprinkle(ref, b"RYKMSWN", 1e-5, rng)
    q = genome.soft_mask(q, 0.4, rng)
    q = genome.insert_runs(q, b"N", 20, 50_000, rng)
    # N only in the query: the reference's host RevComp (common/ntcoding.cpp:63-105) silently drops every other
    # IUPAC letter, which shifts its minus-strand seeds against its own device-side reverse complement -- outside
    # defined behaviour, so not a parity input (tests/harness.py keeps such queries to the plus strand)
    q = genome.sprinkle(q, b"N", 1e-5, rng)
    return ref, q


def extra_workloads(be, args, pool, nthreads):
    """BASELINE configs[0] (E. coli-size self-alignment and its 40 %-diverged variant) and configs[1]
    (the round-1 headline) through the resident leg, a few steps each."""
    import torch
    from segalign_b200.backend import shape_pattern
    span = len(shape_pattern(SEED_SHAPE))
    out = {}

    def run(name, ref, query, steps, note, warm=True, transition=True):
        be.ClearQuery(0)
        be.ClearRef()
        be.SendRefWriteRequest(ref, 0, ref.size)
        be.GenerateSeedPosTable(ref, 0, ref.size, 1)
        be.SendQueryWriteRequest(query, 0, query.size, 0)
        units = genome.chunk_list(query.size, span, "both")

        def work(u):
            rev, j0, j1 = units[u]
            res, ns = be.SeedAndFilterRange(j0, j1, transition, bool(rev), 0)
            return res.size - 1
        if warm:
            hs = sum(pool.map(work, range(len(units))))
        torch.cuda.synchronize()
        be.reset_stats()
        t0 = time.perf_counter()
        for _ in range(steps):
            hs = sum(pool.map(work, range(len(units))))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        st = be.stats()
        out[name] = {"value": round(query.size / dt / 1e9, 5), "unit": "Gbp/s", "ms_per_step": round(dt * 1e3, 3),
                     "steps": steps, "ref_bp": int(ref.size), "query_bp": int(query.size), "hsps": int(hs),
                     "hits_per_step": int(st["hits"] // steps), "workload": note}

    rng = np.random.default_rng(20261017)
    g = genome.random_genome(4_641_652, rng)
    run("configs0_ecoli_mut40", g, genome.mutate(g, 0.40, rng), 3,
        "BASELINE configs[0] throughput variant: 4 641 652 bp genome vs its 40 %-diverged copy (SURVEY 8d config 1)")
    run("configs0_ecoli_self", g, g.copy(), 1,
        "BASELINE configs[0]: 4 641 652 bp genome against itself (one 4.6 Mb main-diagonal HSP that every one of "
        "its seed hits extends to); one cold step", warm=False)
    if args.workload != "ce11":
        a = argparse.Namespace(**vars(args))
        a.workload, a.ref_mb, a.query_mb = "ce11", None, None
        w1 = Workload(a, 0)
        run("configs1_ce11_scale", w1.ref, w1.query, 3, w1.label)
    if not args.no_chr1:
        # BASELINE configs[3] (hg38 chr1 vs mm39 chr1, --notransition --ambiguous=iupac) at its block size: one
        # 248 Mb record, half of it soft-masked, an 18 Mb and forty 50 kb runs of N, IUPAC letters at 1e-5; the query
        # is a 100 Mb slice of its 40 %-diverged copy with its own masking and N runs.  Other scoring options than
        # the main legs: the processor is set up again (the last thing this function does).
        # A secondary workload must not cost the headline line: anything but a parity failure (AssertionError:
        # records differ from the reference kernels') is reported in place of the number.
        name = "configs3_chr1_scale_notransition_iupac"
        try:
            ref, q = make_chr1_pair()
            be.ShutdownProcessor()
            be.InitializeInterface(1, first_device=int(os.environ.get("LOCAL_RANK", "0")))
            be.GenerateShapePos(SEED_SHAPE)
            be.InitializeProcessor(False, genome.DEFAULT_WGA_CHUNK, span, iupac_matrix(), XDROP, HSPTHRESH, False)
            run(name, ref, q, 3,
                "BASELINE configs[3] flags at chr1 scale, synthetic: 248 Mb reference record (50 % soft-masked, 18 Mb + 40 x 50 kb "
                "of N, IUPAC letters) x 100 Mb query slice (40 % diverged), --notransition --ambiguous=iupac; 1 GPU",
                transition=False)
            if not args.no_reference_gpu:
                r = reference_gpu_compare(be, ref, q, span, args, False, iupac_matrix(), args.reference_gpu_mb)
                out[name]["reference_gpu"] = {
                    k: r[k] for k in ("seconds", "ours_seconds", "speedup", "identical", "calls", "hits", "hsps",
                                      "query_slice_bp", "unavailable") if k in r}
        except AssertionError:
            raise
        except Exception as e:  # noqa: BLE001 -- host memory, a missing runner, a time-out
            out.setdefault(name, {})["error"] = "%s: %s" % (type(e).__name__, str(e)[:300])
    return out


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


def lastz_run(ref_fa: Path, pieces) -> float:
    """One single-threaded LASTZ process per query file, all at once (the reference's own wrapper
    parallelises LASTZ this way, scripts/run_segalign:115).  Returns wall seconds."""
    t0 = time.perf_counter()
    procs = []
    for p in pieces:
        procs.append(subprocess.Popen([str(LASTZ), f"{ref_fa}[multiple]", str(p), *LASTZ_ARGS],
                                      stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
    for p in procs:
        p.wait()
    return time.perf_counter() - t0


def lastz_procs(ref_bases: int) -> int:
    """Processes LASTZ can run side by side: one per core, capped by host memory (each process holds the
    target and its own seed position table, ~6 bytes per target base)."""
    cores = os.cpu_count() or 1
    try:
        avail = 0
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                avail = int(line.split()[1]) * 1024
        if avail:
            cores = max(1, min(cores, int(avail * 0.7 / max(1, 8 * ref_bases))))
    except OSError:
        pass
    return cores


def lastz_baseline(wl, budget_s=20.0, cores=None, pieces_per_proc=1):
    """Bounded sample: every core aligns its own query piece(s) against the full reference block.
    LASTZ's per-process target loading + seed-position-table build is timed separately with a
    100-base query and subtracted (the metric counts query throughput with the index resident, as
    for the GPU arm whose table build is reported under setup_ms)."""
    if not LASTZ.exists():
        return {"value": None, "unit": "Gbp/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/lastz missing"}
    ref_chroms, query = wl.ref_chroms, wl.query
    ref_bases = sum(c.size for c in ref_chroms)
    cores = cores or lastz_procs(ref_bases)
    tmp = Path(tempfile.mkdtemp(prefix="sa_lastz_"))
    ref_fa = tmp / "ref.fa"
    write_fasta(ref_fa, ref_chroms, "r")
    rng = np.random.default_rng(7)
    tiny = [tmp / f"tiny{i}.fa" for i in range(cores)]
    for t in tiny:
        write_fasta(t, [genome.random_genome(100, rng)], "t")
    t_setup = lastz_run(ref_fa, tiny)
    # size the sample for ~budget_s of work per core from LASTZ's measured cost per seed hit on these
    # hosts (~0.3 us: 2.1e-5 Gbp/s per core at 155 hits per query base, BENCH_r01) -- no calibration
    # run: at 500 Mb every extra LASTZ start costs its target load + table build again
    hits_per_base = 2 * 13 * ref_bases / 4 ** 12
    clean = lambda a: np.where(a == ord("&"), ord("N"), a).astype(np.uint8)  # noqa: E731
    total_len = int(min(query.size // cores, max(2_000, budget_s * 3.2e6 / hits_per_base)))
    piece_len = max(500, total_len // pieces_per_proc)
    starts = (np.arange(cores) * (query.size // cores)).astype(np.int64)
    pieces = []
    for i, s in enumerate(starts):
        p = tmp / f"piece{i}.fa"
        write_fasta(p, [clean(query[s + k * piece_len:s + (k + 1) * piece_len]) for k in range(pieces_per_proc)], "q")
        pieces.append(p)
    t_run = lastz_run(ref_fa, pieces)
    for f in tmp.glob("*"):
        f.unlink()
    tmp.rmdir()
    subtracted = t_run - t_setup > 0.25 * t_run
    work = t_run - t_setup if subtracted else t_run
    bases = piece_len * pieces_per_proc * cores
    return {"value": round(bases / work / 1e9, 7), "unit": "Gbp/s", "cores": cores, "kind": "reference",
            "sample": "LASTZ 1.04.17 (reference submodule): %d processes x %d query piece(s) of %d bp vs the full %d bp "
                      "reference; %.1f s wall %s %.1f s per-process target load + table build"
                      % (cores, pieces_per_proc, piece_len, ref_bases, t_run,
                         "minus" if subtracted else "(not reduced by the)", t_setup),
            "host_cores": os.cpu_count(),
            "per_core": round(bases / work / 1e9 / cores, 8), "wall_s": round(t_run, 2), "setup_s": round(t_setup, 2),
            "work_s": round(work, 2)}


def run_reference(args):
    """--impl reference: LASTZ on the host cores, same config / metric / unit as our arm.  Every LASTZ
    process has to load the reference block and build its own seed position table before it sees a
    query base (tens of seconds at 500 Mb), so the W+K steps are W+K query pieces handed to ONE process
    per core in a single run; the load + build time (measured with a 100-base query) is subtracted and
    the remaining work is divided evenly over the pieces (LASTZ is deterministic CPU code: no warm-up
    effect to separate)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = Workload(args, 0)
    n_total = args.steps + args.warmup
    t0 = time.perf_counter()
    last = lastz_baseline(wl, budget_s=max(20.0, args.cpu_budget * 3), pieces_per_proc=n_total)
    wall = time.perf_counter() - t0
    v = float(last["value"] or 0.0)
    line = {"impl": "reference", "metric": METRIC,
            "value": round(v, 7), "unit": "Gbp/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(last.get("work_s", wall) / n_total * 1e3, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": wl.label + "; each step = bounded sample: " + last["sample"], "baseline_config": wl.config_id},
            "cpu_baseline": {**last, "value": round(v, 7)},
            "e2e": {"value": round(v, 7), "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": round(wall, 1)}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="syn500", choices=["syn500", "ce11"],
                    help="syn500 = BASELINE configs[2] (default), ce11 = configs[1]")
    ap.add_argument("--no-chr1", action="store_true", help="skip the configs[3]-scale secondary workload")
    ap.add_argument("--strong", action="store_true",
                    help="under torchrun: all ranks share ONE query block and split its calls (strong scaling); "
                         "rank 0 then checks the union against running every call itself")
    ap.add_argument("--inproc", action="store_true", help="one process, --gpus N GPUs in the library's pool (not under torchrun)")
    ap.add_argument("--ref-mb", type=float, default=None, help="scale the reference block (testing only)")
    ap.add_argument("--query-mb", type=float, default=None, help="query block per GPU per step (syn500: default 100)")
    ap.add_argument("--host-threads", type=int, default=0, help="0 = auto: min(16 per GPU, cores / ranks)")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true")
    ap.add_argument("--reference-gpu-mb", type=float, default=5.0, help="query slice handed to the reference's own kernels")
    ap.add_argument("--reference-gpu-timeout", type=int, default=240)
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workloads (configs[0], configs[1])")
    ap.add_argument("--vector-steps", type=int, default=3)
    ap.add_argument("--roofline-launches", type=int, default=160)
    ap.add_argument("--acct-launches", type=int, default=40)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
