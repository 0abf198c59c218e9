"""Large-input parity + timing against the UNMODIFIED reference backend, on a GPU box.

    gpurun -- 'python tests/golden/check_large.py > gpurun_out/check_large.json'

For each case: oracle/_ref/oracle_runner (the reference's own CUDA kernels rebuilt for sm_100a,
SURVEY 8d "reference GPU timing") and the new backend run the same SeedAndFilter calls; every
call's records must be byte-identical; the time both spend inside SeedAndFilter is reported.
The same cases run as `-m gpu` tests (tests/test_live_reference_gpu.py); this script adds the timing
columns and runs the self-alignment at full E. coli size (~270 s of reference kernel time).
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

from segalign_b200 import genome  # noqa: E402
from tests import harness as H  # noqa: E402


os.environ.setdefault("SEGALIGN_LIVE_FULL", "1")   # the self-alignment at E. coli size (the -m gpu test uses 1.2 Mb)
from tests import test_live_reference_gpu as L  # noqa: E402  (generators + case list live with the test)

CASES = L.LIVE_CASES


def main():
    from segalign_b200.backend import Backend
    only = sys.argv[1:]
    out = {}
    work = Path(tempfile.mkdtemp(prefix="sa_large_"))
    for case in CASES:
        if only and case.name not in only:
            continue
        ref, query = case.inputs()
        t0 = time.time()
        dump = H.run_runner(H.ORACLE_RUNNER, case, work)
        t_ref_total = time.time() - t0
        want = []
        for rev, cs, ce, ns, tot, nh, segs in dump.calls:
            res = np.zeros(segs.size + 1, dtype=H.SEGMENT_DTYPE)
            res[0]["len"], res[0]["score"] = tot, np.uint32(nh).view(np.int32)
            res[1:] = segs
            want.append((rev, cs, ce, ns, res))
        row = {"ref_bp": int(ref.size), "query_bp": int(query.size), "calls": len(want),
               "hits": int(dump.counters[1]), "hsps": int(dump.counters[2]),
               "reference_seed_and_filter_s": round(float(dump.times[4]), 3),
               "reference_table_build_s": round(float(dump.times[1]), 3),
               "reference_host_seedgen_s": round(float(dump.times[3]), 3)}
        for mode, dev in (("vector_abi", False), ("device_seeding", True)):
            be = Backend()
            be.InitializeInterface(1)
            span, _ = H.setup_backend(be, case, ref, query)
            from segalign_b200.backend import shape_pattern
            pattern = shape_pattern(case.seed_shape)
            q_rc = genome.revcomp_ascii(query)
            got, t_calls = [], 0.0
            for rev, j0, j1 in H.chunk_calls(case, query.size, span):
                if dev:
                    t1 = time.perf_counter()
                    res, ns = be.SeedAndFilterRange(j0, j1, case.transition, bool(rev), 0)
                    t_calls += time.perf_counter() - t1
                else:
                    seeds = be.host_chunk_seeds(q_rc if rev else query, j0, j1, case.transition)
                    ns = seeds.size
                    if ns == 0:
                        continue
                    t1 = time.perf_counter()
                    res = be.SeedAndFilter(seeds, bool(rev), 0)
                    t_calls += time.perf_counter() - t1
                if ns:
                    got.append((rev, j0, j1, ns, res))
            be.ClearQuery(0); be.ClearRef(); be.ShutdownProcessor()
            try:
                H.assert_calls_equal(got, want, f"{case.name}/{mode}")
                row[mode] = "bit-identical"
            except AssertionError as e:
                row[mode] = f"MISMATCH: {e}"
            row[mode + "_seed_and_filter_s"] = round(t_calls, 3)
        row["speedup_vs_reference_kernels_vector_abi"] = round(row["reference_seed_and_filter_s"] / max(1e-9, row["vector_abi_seed_and_filter_s"]), 1)
        out[case.name] = row
        print(json.dumps({case.name: row}), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
