"""Host-side sharding logic on CPU: index arithmetic and the world_size-2 gloo gather."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def test_shard_ranges_partition_the_batch():
    from ninterp_b200.sharding import shard_range, shard_sizes

    for n in (0, 1, 7, 8, 1000, 10**9 + 3):
        for w in (1, 2, 3, 4, 8):
            rs = [shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            sizes = shard_sizes(n, w)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, n, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle_binding as O
        from ninterp_b200.sharding import gather_results, shard_columns

        # every rank evaluates its shard (here with the CPU oracle standing in for the GPU call: the test is
        # about the host-side plumbing) and the gathered result must equal the unsharded evaluation
        rng = np.random.default_rng(5)
        axes = [np.cumsum(rng.uniform(0.5, 1.5, 9)) for _ in range(3)]
        values = rng.uniform(-1, 1, (9, 9, 9))
        cols = [rng.uniform(a[0] - 0.2, a[-1] + 0.2, n) for a in axes]
        o = O.Oracle(axes, values, O.LINEAR, O.ERROR)
        mine = shard_columns(cols, rank, world)
        out, st = o.eval(mine)
        full_out = gather_results(torch.from_numpy(out), n)
        full_st = gather_results(torch.from_numpy(st), n)
        want, st_want = o.eval(cols)
        ok = full_out.numpy().tobytes() == want.tobytes() and np.array_equal(full_st.numpy(), st_want)
        # error counters are reduced on the host side
        cnt = torch.tensor([int((st == 1).sum())])
        dist.all_reduce(cnt)
        ok = ok and int(cnt.item()) == int((st_want == 1).sum())
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [1001, 4096])
def test_gather_results_world2_gloo(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
