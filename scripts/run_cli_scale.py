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


def digest(folder: Path) -> str:
    """sha256 over (name, bytes) of every file of the output folder, in name order."""
    import hashlib
    h = hashlib.sha256()
    for f in sorted(folder.glob("*")):
        h.update(f.name.encode())
        data = f.read_bytes()
        if f.name == "lastz_commands.txt":   # one line per finished interval, in completion order (as the reference's printer)
            data = b"\n".join(sorted(data.split(b"\n")))
        h.update(data)
    return h.hexdigest()


def main():
    rng = np.random.default_rng(20261017)
    chr1 = "--chr1" in sys.argv[1:]
    if chr1:
        # BASELINE configs[3] at its block size through the C++ driver: bench.py's secondary workload of that name
        # (one 248 Mb record x a 100 Mb query record, --notransition --ambiguous=iupac), one block each
        import bench
        r, q = bench.make_chr1_pair()
        ref, qry = [r], [q]
        flags, block = ["--notransition", "--ambiguous=iupac"], 500_000_000
    else:
        # reference: 6 records of 52 Mb -> blocks close behind the record that passes 100 Mb: 3 blocks of 104 Mb
        ref = [genome.soft_mask(genome.random_genome(52_000_000, rng), 0.3, rng) for _ in range(6)]
        # query: 30 %-diverged copies of 4 of them + their own masking -> 2 blocks of 104 Mb
        qry = [genome.soft_mask(genome.mutate(np.where(c >= 97, c - 32, c).astype(np.uint8), 0.30, rng), 0.3, rng) for c in ref[:4]]
        flags, block = [], 100_000_000
    work = Path(tempfile.mkdtemp(prefix="sa_cli_"))
    fasta(work / "ref.fa", ref, "chrR")
    fasta(work / "query.fa", qry, "chrQ")
    out = work / "out"
    out.mkdir()
    extra_args = [a for a in sys.argv[1:] if a.startswith("--num_gpu")]
    one_gpu_digest = None
    if "--compare-one-gpu" in sys.argv[1:]:
        # the same input on ONE GPU first: every output file must come out byte-identical on the whole pool
        p1 = subprocess.run([str(CLI), str(work / "ref.fa"), str(work / "query.fa"), "/data", f"--out_dir={out}",
                             f"--seq_block_size={block}", "--nogapped", "--num_gpu=1", *flags], capture_output=True, text=True)
        if p1.returncode != 0:
            print(json.dumps({"error": p1.stderr[-500:], "rc": p1.returncode, "run": "one gpu"}))
            return 1
        one_gpu_line = [l for l in p1.stderr.splitlines() if l.startswith("ref blocks")][-1]
        one_gpu_digest = digest(out)
        for f in out.glob("*"):
            f.unlink()
    t0 = time.perf_counter()
    p = subprocess.run([str(CLI), str(work / "ref.fa"), str(work / "query.fa"), "/data", f"--out_dir={out}",
                        f"--seq_block_size={block}", "--nogapped", *flags, *extra_args], capture_output=True, text=True)
    wall = time.perf_counter() - t0
    if p.returncode != 0:
        print(json.dumps({"error": p.stderr[-500:], "rc": p.returncode}))
        return 1
    line = [l for l in p.stderr.splitlines() if l.startswith("ref blocks")][-1]
    nums = [float(x) for x in re.findall(r"[-+]?\d*\.?\d+", line)]
    segs = list(out.glob("*.segments"))
    qbases = sum(c.size for c in qry)
    res = {"what": "segalign_b200_cli (sa_pipeline_run): FASTA -> blocks -> every reference block x every query block -> tmp*.segments",
           "ref_bp": int(sum(c.size for c in ref)), "query_bp": int(qbases), "seq_block_size": block, "flags": flags,
           "summary_line": line, "wall_s_incl_fasta_read": round(wall, 2),
           "ref_blocks": int(nums[0]), "query_blocks": int(nums[1]), "intervals": int(nums[2]), "calls": int(nums[3]),
           "hits": int(nums[5]), "hsps": int(nums[6]), "segment_files": len(segs), "driver_seconds": nums[8],
           "ms_ref_upload_encode_total": nums[9], "ms_seed_pos_tables_total": nums[10], "ms_query_upload_encode_total": nums[11],
           "seconds_read_input": nums[12], "seconds_device_init": nums[13], "seconds_align": nums[14],
           "gbp_query_x_ref_blocks_per_s": round(qbases * nums[0] / nums[8] / 1e9, 4),
           "gbp_query_x_ref_blocks_per_s_alignment_only": round(qbases * nums[0] / max(1e-9, nums[14]) / 1e9, 4),
           "segments_bytes": sum(f.stat().st_size for f in segs)}
    gl = [l for l in p.stderr.splitlines() if "GPU" in l]
    if gl:
        res["gpu_line"] = gl[0][:200]
    if one_gpu_digest is not None:
        res["one_gpu"] = {"summary_line": one_gpu_line, "output_files_identical_to_pool_run": digest(out) == one_gpu_digest}
        if not res["one_gpu"]["output_files_identical_to_pool_run"]:
            print(json.dumps(res))
            return 2
    print(json.dumps(res))
    for f in list(out.glob("*")) + list(work.glob("*.fa")):
        f.unlink()
    out.rmdir(); work.rmdir()
    return 0


if __name__ == "__main__":
    sys.exit(main())
