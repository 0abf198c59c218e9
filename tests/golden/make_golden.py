"""Generates the golden fixtures tests/golden/*.npz by running the UNMODIFIED reference backend.

Runs on a GPU box (the reference backend is CUDA-only):

    gpurun -- 'python tests/golden/make_golden.py --out gpurun_out/golden'
    cp gpurun_out/golden/*.npz tests/golden/

For every case of tests/harness.py:CASES it writes the case file, runs
oracle/_ref/oracle_runner (reference objects compiled from /root/reference by oracle/Makefile,
plus oracle/ref_driver.cpp) with --check-seeder (the reference's own seeder_body must return the
same records as the driver's chunk loop) and stores every SeedAndFilter return value.  It then
cross-checks the CPU restatement (oracle/sa_oracle.c) and, if --backend is given, the new CUDA
backend against those dumps and prints one line per case.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

from tests import harness as H  # noqa: E402


def dump_as_calls(dump):
    out = []
    for rev, cs, ce, ns, tot, nh, segs in dump.calls:
        res = np.zeros(segs.size + 1, dtype=H.SEGMENT_DTYPE)
        res[0]["len"], res[0]["score"] = tot, np.uint32(nh).view(np.int32)
        res[1:] = segs
        out.append((rev, cs, ce, ns, res))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "golden"))
    ap.add_argument("--backend", action="store_true", help="also check the new CUDA backend")
    ap.add_argument("--cases", nargs="*", default=None)
    args = ap.parse_args()
    out = Path(args.out)
    out.mkdir(parents=True, exist_ok=True)
    H.GOLDEN_DIR = out
    work = out / "_tmp"
    report = {}
    be = None
    if args.backend:
        from segalign_b200.backend import Backend
        be = Backend()
    for case in H.CASES:
        if args.cases and case.name not in args.cases:
            continue
        ref, query = case.inputs()
        t0 = time.time()
        try:
            dump = H.run_runner(H.ORACLE_RUNNER, case, work, extra=("--check-seeder", "--dump-table"))
        except Exception as e:  # noqa: BLE001
            report[case.name] = {"reference": f"FAILED: {e} {getattr(e, 'stderr', b'')[-400:]}"}
            print(case.name, report[case.name], flush=True)
            continue
        t_ref = time.time() - t0
        H.save_golden(case, dump, ref, query)
        want = dump_as_calls(dump)
        row = {"calls": len(want), "hits": int(dump.counters[1]), "hsps": int(dump.counters[2]),
               "max_hits_device": int(dump.counters[3]), "ref_s": round(t_ref, 2)}
        try:
            got = H.run_cpu_oracle(case, ref, query, max_hits_device=int(dump.counters[3]))
            H.assert_calls_equal(got, want, "cpu-oracle vs reference")
            row["cpu_oracle"] = "OK"
        except AssertionError as e:
            row["cpu_oracle"] = f"MISMATCH: {e}"
        # table: same index, same bucket multisets
        from oracle import sa_oracle_py as sao
        idx, pos = dump.table
        tab = sao.Table(sao.Shape(case.seed_shape), ref, ref.size, case.step)
        ok = np.array_equal(idx, tab.index) and pos.size == tab.pos.size
        if ok and pos.size:
            starts = np.concatenate([[0], idx[:-1]]).astype(np.int64)
            bucket = np.repeat(np.arange(idx.size, dtype=np.int64), (idx.astype(np.int64) - starts))
            a = np.lexsort((pos, bucket)); b = np.lexsort((tab.pos, bucket))
            ok = np.array_equal(pos[a], tab.pos[b])
        row["cpu_table"] = "OK" if ok else "MISMATCH"
        if be is not None:
            for dev_seed in (False, True):
                key = "backend_range" if dev_seed else "backend"
                try:
                    be.InitializeInterface(1)
                    got = H.run_backend(be, case, ref, query, device_seeding=dev_seed)
                    H.assert_calls_equal(got, want, f"{key} vs reference")
                    row[key] = "OK"
                except Exception as e:  # noqa: BLE001
                    row[key] = f"MISMATCH: {type(e).__name__}: {e}"
        report[case.name] = row
        print(case.name, json.dumps(row), flush=True)
    (out / "report.json").write_text(json.dumps(report, indent=1))
    for f in work.glob("*"):
        f.unlink()
    work.rmdir()
    bad = [k for k, v in report.items() if any(isinstance(x, str) and x.startswith("MISMATCH") for x in v.values())]
    print("golden: %d cases, %d with mismatches %s" % (len(report), len(bad), bad))
    return 0


if __name__ == "__main__":
    sys.exit(main())
