"""CPU-only, world_size 2 over gloo: the multi-GPU path is a batch-slice partition with no data-path collective
(gpu_ntt_b200/multigpu.py); the ranks only agree on timing.  Checks that the slices tile the batch exactly, respect
RNS modulus groups, that each rank would transform exactly its own polynomials (verified here with the oracle on the
slice: the concatenation of per-rank results equals the single-rank result), and the max/sum aggregation."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from gpu_ntt_b200.multigpu import aggregate_rate, batch_slice  # noqa: E402


def test_slices_tile_the_batch():
    for world in (1, 2, 3, 4, 8):
        for batch, mc in ((8192, 0), (1024, 0), (7, 0), (24, 3), (64, 4), (6, 3)):
            cuts = [batch_slice(r, world, batch, mc) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == batch
            for (a, b), (c, d) in zip(cuts, cuts[1:]):
                assert b == c and a <= b
            if mc:
                assert all(a % mc == 0 and b % mc == 0 for a, b in cuts)
    assert batch_slice(3, 8, 8192) == (3072, 4096)          # config C5: 1024 polynomials per GPU
    with pytest.raises(ValueError):
        batch_slice(0, 2, 7, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    logn, batch = 8, 12
    P = O.merge_params(logn, O.X_N_minus, 64)
    x = O.example_input(P.modulus, batch << logn, seed=5).reshape(batch, -1)     # the one global stream of polynomials
    lo, hi = batch_slice(rank, world, batch)
    y_local = O.merge_ntt(np.ascontiguousarray(x[lo:hi]), P)                     # what this rank's GPU would produce
    # no data-path collective exists; for the check only, gather the slices and compare with the unsharded transform
    parts = [None] * world
    dist.all_gather_object(parts, (lo, hi, y_local))
    rate, ms, units = aggregate_rate(hi - lo, 10.0 * (rank + 1), dist)
    if rank == 0:
        parts.sort(key=lambda t: t[0])
        whole = np.concatenate([p[2] for p in parts])
        ok = bool((whole == O.merge_ntt(x, P)).all()) and parts[0][0] == 0 and parts[-1][1] == batch
        q.put((ok, rate, ms, units))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, rate, ms, units = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
    assert units == 12 and ms == 20.0 and abs(rate - 12 / 0.020) < 1e-6     # sum of units / max of times
