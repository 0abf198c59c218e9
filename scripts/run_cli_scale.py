#!/usr/bin/env python
"""BASELINE configs[4] at ~1/10 scale through the whole-genome driver (segalign_b200_cli): several
reference blocks x several query blocks of >= 100 Mb each, from FASTA files to tmp*.segments + LASTZ
command lines.  Reports wall time and the per-block upload / table-build milliseconds.

    gpurun -- 'python scripts/run_cli_scale.py > gpurun_out/cli_scale.json'
"""
import json
import re
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from segalign_b200 import genome  # noqa: E402
from segalign_b200.build import CLI  # noqa: E402


def fasta(path, chroms, prefix):
    with open(path, "wb") as f:
        for i, c in enumerate(chroms):
            f.write(b">%s%d\n" % (prefix.encode(), i))
            f.write(c.tobytes())
            f.write(b"\n")


def main():
    rng = np.random.default_rng(20261017)
    # reference: 6 records of 52 Mb -> blocks close behind the record that passes 100 Mb: 3 blocks of 104 Mb
    ref = [genome.soft_mask(genome.random_genome(52_000_000, rng), 0.3, rng) for _ in range(6)]
    # query: 30 %-diverged copies of 4 of them + their own masking -> 2 blocks of 104 Mb
    qry = [genome.soft_mask(genome.mutate(np.where(c >= 97, c - 32, c).astype(np.uint8), 0.30, rng), 0.3, rng) for c in ref[:4]]
    work = Path(tempfile.mkdtemp(prefix="sa_cli_"))
    fasta(work / "ref.fa", ref, "chrR")
    fasta(work / "query.fa", qry, "chrQ")
    out = work / "out"
    out.mkdir()
    t0 = time.perf_counter()
    p = subprocess.run([str(CLI), str(work / "ref.fa"), str(work / "query.fa"), "/data", f"--out_dir={out}",
                        "--seq_block_size=100000000", "--nogapped"], capture_output=True, text=True)
    wall = time.perf_counter() - t0
    if p.returncode != 0:
        print(json.dumps({"error": p.stderr[-500:], "rc": p.returncode}))
        return 1
    line = [l for l in p.stderr.splitlines() if l.startswith("ref blocks")][-1]
    nums = [float(x) for x in re.findall(r"[-+]?\d*\.?\d+", line)]
    segs = list(out.glob("*.segments"))
    qbases = sum(c.size for c in qry)
    res = {"what": "segalign_b200_cli (sa_pipeline_run): FASTA -> blocks -> every reference block x every query block -> tmp*.segments",
           "ref_bp": int(sum(c.size for c in ref)), "query_bp": int(qbases), "seq_block_size": 100_000_000,
           "summary_line": line, "wall_s_incl_fasta_read": round(wall, 2),
           "ref_blocks": int(nums[0]), "query_blocks": int(nums[1]), "intervals": int(nums[2]), "calls": int(nums[3]),
           "hits": int(nums[5]), "hsps": int(nums[6]), "segment_files": len(segs), "driver_seconds": nums[8],
           "ms_ref_upload_encode_total": nums[9], "ms_seed_pos_tables_total": nums[10], "ms_query_upload_encode_total": nums[11],
           "gbp_query_x_ref_blocks_per_s": round(qbases * nums[0] / nums[8] / 1e9, 4),
           "segments_bytes": sum(f.stat().st_size for f in segs)}
    print(json.dumps(res))
    for f in list(out.glob("*")) + list(work.glob("*.fa")):
        f.unlink()
    out.rmdir(); work.rmdir()
    return 0


if __name__ == "__main__":
    sys.exit(main())
