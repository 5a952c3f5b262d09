"""The N>1 path on CPU: world_size-2 gloo process group, contiguous shards, MAX-over-ranks timing, host gather.

The GPU evaluator cannot run here, so the per-shard evaluation is done by the oracle (the checker standing in
for the device) -- what is under test is the host-side plumbing the bench and the sharded API use."""
import os
import sys

import numpy as np
import pytest

from symbanafis_b200.sharding import shard_bounds, shard_columns

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.mark.parametrize("n,world", [(0, 1), (1, 2), (7, 2), (100_000_000, 8), (12345, 4), (3, 8)])
def test_shards_partition_the_points(n, world):
    prev = 0
    for r in range(world):
        b, e = shard_bounds(n, world, r)
        assert b == prev and e >= b
        prev = e
    assert prev == n
    sizes = [shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world)]
    assert max(sizes) - min(sizes) <= 1


def test_shard_columns_broadcast_rules():
    full = np.arange(10.0)
    short = np.arange(4.0)          # broadcast column: element 3 repeats from point 4 on
    c = shard_columns([full, short], 10, 2, 0)
    assert c[0].tolist() == [0, 1, 2, 3, 4] and c[1].tolist() == [0, 1, 2, 3]
    c = shard_columns([full, short], 10, 2, 1)
    assert c[0].tolist() == [5, 6, 7, 8, 9] and c[1].tolist() == [3]
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _worker(rank: int, world: int, port: int, out_dir: str):
    sys.path[:0] = [ROOT, HERE]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from oracle import OracleProgram
    from symbanafis_b200 import workloads
    from symbanafis_b200.sharding import eval_sharded, gather_outputs, max_over_ranks
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = workloads.get("aizawa")
        n = 10_001  # odd: ragged shards
        cols = w.make_inputs(n, 7)  # every rank builds the same global columns and evaluates only its shard
        op = OracleProgram.from_program(w.program)
        local = eval_sharded(lambda c, m: op.eval_batch(c, m), cols, n, world, rank)
        dist.barrier()
        t = max_over_ranks([float(rank + 1), 10.0 - rank])
        full = gather_outputs(local, n)
        if rank == 0:
            ref = op.eval_batch(cols, n)
            ok = all(np.array_equal(a.view(np.uint64), b.view(np.uint64)) for a, b in zip(full, ref))
            with open(os.path.join(out_dir, "result.txt"), "w") as f:
                f.write(f"{int(ok)} {t[0]} {t[1]}\n")
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo(tmp_path):
    import socket

    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ok, t0, t1 = open(tmp_path / "result.txt").read().split()
    assert ok == "1"                        # gathered shards == the unsharded evaluation, bit for bit
    assert float(t0) == 2.0 and float(t1) == 10.0  # MAX over ranks
