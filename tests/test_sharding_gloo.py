"""N>1 host path on CPU: two gloo ranks shard the SeedAndFilter units of one query block, run
them through the checker (the CPU oracle stands in for the GPU here -- this is a test of the
sharding/merge logic, not of the product path) and the merged result must equal the 1-rank run
('GPU assignment does not affect any result' is the multi-GPU test)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import sa_oracle_py as sao
    from segalign_b200 import sharding
    from tests import harness as H
    case = H.CASES_BY_NAME["diverged_chunked"]
    ref, query = case.inputs()
    shape = sao.Shape(case.seed_shape)
    table = sao.Table(shape, ref, ref.size, case.step)
    ref_enc = sao.encode(ref)
    q_fwd, q_rc = sao.encode_rc(query)
    q_rc_ascii = sao.revcomp_ascii(query)
    params = sao.make_params(H.matrix_for(case), case.xdrop, case.hspthresh, case.noentropy, shape.span, 1 << 30)
    units = H.chunk_calls(case, query.size, shape.span)
    mine = sharding.shard_units(len(units), rank, world)
    local, n_bases = [], 0
    for u in mine:
        rev, j0, j1 = units[u]
        seeds = shape.chunk_seeds(q_rc_ascii if rev else query, j0, j1, case.transition)
        res = sao.seed_and_filter(params, table, ref_enc, q_rc if rev else q_fwd, seeds)
        local.append((u, res[1:].copy()))
        n_bases += j1 - j0
    # the only cross-rank traffic: a scalar reduction for reporting (max time / total units)
    t = torch.tensor([float(n_bases), float(len(local))], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    if rank == 0:
        merged = sharding.merge_ranked(gathered)
        np.savez(Path(out_dir) / "merged.npz", units=np.array([u for u, _ in merged]),
                 segs=np.concatenate([s for _, s in merged]).view(np.uint32).reshape(-1, 4),
                 totals=t.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharding_equals_single_rank(tmp_path):
    from tests import harness as H
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    z = np.load(tmp_path / "merged.npz")
    case = H.CASES_BY_NAME["diverged_chunked"]
    want, _ = H.golden_as_calls(case)
    assert z["units"].tolist() == list(range(len(want)))
    want_segs = np.concatenate([w[4][1:] for w in want])
    got = np.ascontiguousarray(z["segs"]).view(H.SEGMENT_DTYPE).reshape(-1)
    assert np.array_equal(got, want_segs)
    assert z["totals"][1] == len(want)
    _, query = case.inputs()
    assert z["totals"][0] == 2 * (query.size - 19)
