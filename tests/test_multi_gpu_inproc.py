"""The reference's multi-GPU model on hardware: ONE process, InitializeInterface(-1) takes every
visible GPU, the reference block / table / query block are replicated, and concurrent SeedAndFilter
callers are handed to whichever GPU has a free stream (common/seed_filter_interface.cu:49-80,
src/seed_filter.cu:699-708, :798-803).  Results must not depend on which GPU served a call
(SURVEY A.9): byte-identical to the 1-GPU goldens.  Needs >= 2 visible GPUs
(`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu_inproc.py -m gpu`)."""
from concurrent.futures import ThreadPoolExecutor

import pytest

from segalign_b200 import genome
from tests import harness as H

pytestmark = pytest.mark.gpu


def _device_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("name", ["diverged_chunked", "masked_multichrom", "repeats_entropy"])
@pytest.mark.parametrize("device_seeding", [False, True], ids=["vector", "range"])
def test_inprocess_gpu_pool_matches_single_gpu_golden(built, name, device_seeding, monkeypatch):
    n = _device_count()
    if n < 2:
        pytest.skip(f"{n} GPU visible: the in-process pool needs at least 2")
    from segalign_b200.backend import Backend, shape_pattern
    monkeypatch.setenv("SEGALIGN_B200_STREAMS", "2")
    case = H.CASES_BY_NAME[name]
    want, digest = H.golden_as_calls(case)
    ref, query = case.inputs()
    assert H.inputs_digest(ref, query) == digest
    be = Backend()
    assert be.InitializeInterface(-1) == n          # the reference's "use all GPUs"
    span, _ = H.setup_backend(be, case, ref, query)  # uploads + table build on every GPU of the pool
    pattern = shape_pattern(case.seed_shape)
    q_rc = genome.revcomp_ascii(query)
    calls = H.chunk_calls(case, query.size, span)
    be.reset_stats()

    def work(c):
        rev, j0, j1 = c
        if device_seeding:
            res, ns = be.SeedAndFilterRange(j0, j1, case.transition, bool(rev), 0)
        else:
            seeds = genome.chunk_seeds(q_rc if rev else query, j0, j1, pattern, case.transition)
            ns = seeds.size
            res = be.SeedAndFilter(seeds, bool(rev), 0) if ns else None
        return (rev, j0, j1, ns, res)

    try:
        # several rounds so that every GPU of the pool gets calls whatever the timing
        for _ in range(3):
            with ThreadPoolExecutor(max_workers=2 * n) as pool:
                got = [g for g in pool.map(work, calls) if g[3]]
            H.assert_calls_equal(got, want, f"in-process pool of {n} GPUs vs the 1-GPU reference golden")
        per_gpu = be.gpu_calls()
        assert len(per_gpu) == n and sum(per_gpu) == 3 * len(got)
        assert sum(1 for c in per_gpu if c > 0) >= 2, f"calls were not spread over the pool: {per_gpu}"
    finally:
        be.ClearQuery(0); be.ClearRef(); be.ShutdownProcessor()


def test_table_is_identical_on_every_gpu_of_the_pool(built, monkeypatch):
    """Each GPU builds its own seed position table from the block it received by peer copy: the
    single-GPU results above can only match if block and table are the same everywhere; this checks
    the hit totals per call while forcing every call onto a different GPU in turn."""
    n = _device_count()
    if n < 2:
        pytest.skip(f"{n} GPU visible")
    from segalign_b200.backend import Backend
    monkeypatch.setenv("SEGALIGN_B200_STREAMS", "1")   # one stream per GPU: n concurrent callers = one per GPU
    case = H.CASES_BY_NAME["diverged_default"]
    want, _ = H.golden_as_calls(case)
    ref, query = case.inputs()
    be = Backend()
    be.InitializeInterface(-1)
    span, _ = H.setup_backend(be, case, ref, query)
    rev, j0, j1 = H.chunk_calls(case, query.size, span)[0]
    try:
        with ThreadPoolExecutor(max_workers=n) as pool:
            outs = list(pool.map(lambda _: be.SeedAndFilterRange(j0, j1, case.transition, bool(rev), 0)[0], range(4 * n)))
        for res in outs:
            assert res[0]["score"] == want[0][4][0]["score"] and (res[1:] == want[0][4][1:]).all()
    finally:
        be.ClearQuery(0); be.ClearRef(); be.ShutdownProcessor()
