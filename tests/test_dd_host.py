"""CPU tests of the multi-GPU host logic (world_size 2, gloo): slab split, result merging, thermo finalisation and
the collective plumbing of DDEngine -- everything above the C ABI that does not need a device."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_split_covers_the_grid_like_the_engine():
    from lammps_le_b200.engine_dd import slab_of_cells
    for ncx in (7, 64, 112, 224):
        for world in (1, 2, 3, 4, 8):
            sl = slab_of_cells(ncx, world)
            assert sl[0][0] == 0 and sl[-1][1] == ncx
            assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
            w = [b - a for a, b in sl]
            assert max(w) - min(w) <= 1


def test_merge_csr_and_thermo():
    from lammps_le_b200.engine_dd import finalize_thermo, merge_csr
    # rank 0 owns atoms 0 and 2, rank 1 owns atom 1 and 3
    a = (np.array([0, 2, 2, 3, 3]), np.array([5, 6, 7], dtype=np.int32))
    b = (np.array([0, 0, 1, 1, 3]), np.array([9, 1, 2], dtype=np.int32))
    off, ent = merge_csr([a, b])
    assert off.tolist() == [0, 2, 3, 4, 6] and ent.tolist() == [5, 6, 9, 7, 1, 2]
    s = np.zeros(16)
    s[0], s[1], s[2], s[3:6] = 3.0 * 99, 10.0, 20.0, (1.0, 2.0, 3.0)
    t = finalize_thermo(s, 100, 1000.0, step=7, nbonds=99)
    assert abs(t["temp"] - 1.0) < 1e-12 and abs(t["epair"] - 0.1) < 1e-12 and abs(t["emol"] - 0.2) < 1e-12
    assert abs(t["press"] - (297.0 + 6.0) / 3000.0) < 1e-12 and t["step"] == 7


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from lammps_le_b200 import engine_dd
    r, w, local, group = engine_dd.init_process_group()
    # the collective helpers of DDEngine without a context: owned entries are summed over ranks
    class Stub(engine_dd.DDEngine):
        def __init__(self):
            self.rank, self.world, self.group = r, w, group
    e = Stub()
    n = 10
    owned = np.arange(n) % w == r
    x = np.where(owned[:, None], np.arange(n * 3, dtype=np.float64).reshape(n, 3), 0.0)
    e._allreduce(x)
    ok = bool((x == np.arange(n * 3).reshape(n, 3)).all())
    parts = [None] * w
    dist.all_gather_object(parts, (np.array([0, 1, 1]) if r == 0 else np.array([0, 0, 2]), np.array([4 + r] * (1 + r), dtype=np.int32)), group=group)
    off, ent = engine_dd.merge_csr(parts)
    ok = ok and off.tolist() == [0, 1, 3] and ent.tolist() == [4, 5, 5]
    e.barrier()
    q.put((rank, ok))
    dist.destroy_process_group()


def test_dd_collectives_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
