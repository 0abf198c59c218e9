"""Per-phase GPU time of SeedAndFilter calls on a secondary workload (library CUDA events, serialized calls).
    python scripts/experiments/phase_probe.py chr1|ecoli [calls]"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from segalign_b200 import genome  # noqa: E402
from segalign_b200.backend import Backend, shape_pattern  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "chr1"
n_calls = int(sys.argv[2]) if len(sys.argv) > 2 else 40
span = len(shape_pattern(bench.SEED_SHAPE))
be = Backend()
be.InitializeInterface(1)
be.GenerateShapePos(bench.SEED_SHAPE)
if what == "chr1":
    ref, q = bench.make_chr1_pair(query_n=20_000_000)
    transition, mat = False, bench.iupac_matrix()
else:
    rng = np.random.default_rng(1)
    ref = genome.random_genome(4_641_652, rng)
    q = genome.mutate(ref, 0.40, rng)
    transition, mat = True, bench.default_matrix()
be.InitializeProcessor(transition, genome.DEFAULT_WGA_CHUNK, span, mat, bench.XDROP, bench.HSPTHRESH, False)
be.SendRefWriteRequest(ref, 0, ref.size)
be.GenerateSeedPosTable(ref, 0, ref.size, 1)
be.SendQueryWriteRequest(q, 0, q.size, 0)
units = genome.chunk_list(q.size, span, "both")[:n_calls]
for rev, j0, j1 in units[:4]:
    be.SeedAndFilterRange(j0, j1, transition, bool(rev), 0)
be.set_profiling(True)
be.reset_stats()
import time
t0 = time.perf_counter()
for rev, j0, j1 in units:
    be.SeedAndFilterRange(j0, j1, transition, bool(rev), 0)
dt = time.perf_counter() - t0
st = be.stats()
n = max(1, st["calls"])
print(json.dumps({"workload": what, "calls": n, "wall_ms_per_call": round(dt * 1e3 / n, 4),
                  "per_call": {k: (round(v / n, 4) if k.startswith("ms_") else v // n) for k, v in st.items()
                               if isinstance(v, (int, float))}}))
