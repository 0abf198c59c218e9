"""Golden fixtures of the repeat-masker variant: tests/golden/rm/*.npz.

Runs on a GPU box:

    gpurun -- 'python tests/golden/make_golden_rm.py --out gpurun_out/golden_rm --backend'
    cp gpurun_out/golden_rm/*.npz gpurun_out/golden_rm/report.json tests/golden/rm/

For every case of tests/harness.py:RM_CASES it runs oracle/_ref/rm_oracle_runner -- the reference's
UNMODIFIED repeat_masker_src/seed_filter.cu + common/*.cu compiled by oracle/Makefile, behind
oracle/rm_driver.cpp -- and stores every SeedAndFilter(seeds, rev, ref_start, ref_end) return value.
It then cross-checks the CPU restatement (oracle/sa_oracle.c: sao_rm_*) and, with --backend, the CUDA
backend through both entry points and through the drop-in runner (rm_new_runner)."""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

from tests import harness as H  # noqa: E402


def dump_as_calls(dump):
    import numpy as np
    out = []
    for rev, cs, ce, ns, rs, re_, hdr, segs in dump.calls:
        res = np.zeros(segs.size + 1, dtype=H.SEGMENT_DTYPE)
        res[0] = (hdr[0], hdr[1], hdr[2], np.uint32(hdr[3]).view(np.int32))
        res[1:] = segs
        out.append((rev, cs, ce, ns, rs, re_, res))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "golden_rm"))
    ap.add_argument("--backend", action="store_true")
    args = ap.parse_args()
    out = Path(args.out)
    out.mkdir(parents=True, exist_ok=True)
    H.RM_GOLDEN_DIR = out
    work = out / "_tmp"
    report = {}
    for case, prop in H.RM_CASES:
        seq, _ = case.inputs()
        t0 = time.time()
        try:
            dump = H.run_rm_runner(H.RM_ORACLE_RUNNER, case, prop, work)
        except Exception as e:  # noqa: BLE001
            report[case.name] = {"reference": f"FAILED: {e} {getattr(e, 'stderr', b'')[-400:]}"}
            print(case.name, report[case.name], flush=True)
            continue
        H.save_rm_golden(case, dump, seq)
        want = dump_as_calls(dump)
        row = {"calls": len(want), "hits": int(dump.counters[1]), "hsps": int(dump.counters[2]),
               "max_hits_device": int(dump.counters[3]), "ref_s": round(time.time() - t0, 2)}
        try:
            got = H.run_rm_cpu_oracle(case, prop, seq, max_hits_device=int(dump.counters[3]))
            H.assert_rm_calls_equal(got, want, "cpu-oracle vs reference")
            row["cpu_oracle"] = "OK"
        except AssertionError as e:
            row["cpu_oracle"] = f"MISMATCH: {e}"
        if args.backend:
            from segalign_b200.backend import Backend
            for dev_seed in (False, True):
                key = "backend_range" if dev_seed else "backend"
                try:
                    be = Backend()
                    be.InitializeInterface(1)
                    got = H.run_rm_backend(be, case, prop, seq, device_seeding=dev_seed)
                    H.assert_rm_calls_equal(got, want, f"{key} vs reference")
                    row[key] = "OK"
                except Exception as e:  # noqa: BLE001
                    row[key] = f"MISMATCH: {type(e).__name__}: {e}"
            try:
                d2 = H.run_rm_runner(H.RM_NEW_RUNNER, case, prop, work)
                H.assert_rm_calls_equal(dump_as_calls(d2), want, "drop-in runner vs reference")
                row["dropin_runner"] = "OK"
            except Exception as e:  # noqa: BLE001
                row["dropin_runner"] = f"MISMATCH: {type(e).__name__}: {e}"
        report[case.name] = row
        print(case.name, json.dumps(row), flush=True)
    (out / "report.json").write_text(json.dumps(report, indent=1))
    if work.exists():
        for f in work.glob("*"):
            f.unlink()
        work.rmdir()
    bad = [k for k, v in report.items() if any(isinstance(x, str) and (x.startswith("MISMATCH") or x.startswith("FAILED")) for x in v.values())]
    print("golden rm: %d cases, %d with mismatches %s" % (len(report), len(bad), bad))
    return 0


if __name__ == "__main__":
    sys.exit(main())
