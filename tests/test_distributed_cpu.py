"""world_size-2 gloo tests (CPU) of the only multi-rank logic: contiguous shot-range sharding and the single
all-reduce of the counter vector."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_shard_range_partitions_exactly():
    from qldpc_b200.experiments import shard_range
    for n in (0, 1, 7, 10_000_000, 10**9 + 3):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (f0, c0), (f1, _) in zip(parts, parts[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def _worker(rank, world, port, nshots, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from qldpc_b200 import experiments as X, _lib
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class FakeCode:            # counters that depend only on the global shot ids, like the Philox-keyed device sweep
        def mc_sweep(self, p, count, seed=0, first_shot=0, draws=1, **kw):
            ids = np.arange(first_shot, first_shot + count, dtype=np.int64)
            c = dict.fromkeys(_lib.COUNTER_NAMES, 0)
            c["shots"] = int(count)
            c["logical"] = int(((ids * 2654435761 + seed) % 97 < 5).sum())
            c["bp_failed"] = int(((ids * 40503 + 7 * seed) % 89 < 9).sum())
            c["iter_sum"] = int((ids % 13).sum())
            return c

    got = X.mc_point(FakeCode(), 0.05, nshots, seed=3)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([got[k] for k in _lib.COUNTER_NAMES], np.int64))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_counters_are_shard_invariant_gloo(tmp_path, world):
    import socket
    import torch.multiprocessing as mp
    from qldpc_b200 import _lib
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    nshots = 100_003
    mp.spawn(_worker, args=(world, port, nshots, str(tmp_path)), nprocs=world, join=True)
    ids = np.arange(nshots, dtype=np.int64)
    want = dict.fromkeys(_lib.COUNTER_NAMES, 0)
    want.update(shots=nshots, logical=int(((ids * 2654435761 + 3) % 97 < 5).sum()),
                bp_failed=int(((ids * 40503 + 21) % 89 < 9).sum()), iter_sum=int((ids % 13).sum()))
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"r{r}.npy"))
        assert got.tolist() == [want[k] for k in _lib.COUNTER_NAMES]
