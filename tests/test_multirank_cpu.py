"""world_size-2 gloo test (CPU): columns shard across ranks with no hot-path communication; the only
collective is the all-reduce of the six budget scalars.  The sharded run must reproduce the
single-rank per-column outputs exactly and the global sums within 1e-13 relative."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _budget(o, ch, r):
    g = o.params.gravit
    w = np.zeros(6)
    for c in range(ch.nchunks):
        n = int(ch.ncol[c])
        w[0] += (ch.pdel[c][:, :n] / g * r["ptend_q"][c][:, :n]).sum()
        w[1] += 1000.0 * (r["prec"][c][:n] + r["rliq"][c][:n]).sum()
        w[2] += (ch.pdel[c][:, :n] / g * r["ptend_s"][c][:, :n]).sum()
        w[4] += r["lengath"][c]
        w[5] += n
    return w


def _worker(rank, world, port, ncols, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cam_nor_physics_b200 import soundings as S
    from helpers import get_oracle
    o, _, _ = get_oracle("libm", 16, 32)
    import bench                                     # the product's host-side partition of the grid (bench.shard_of)
    col0, per = bench.shard_of(rank, world, ncols, "strong")
    ch = S.make_chunks(per, 32, 16, p_conv=0.5, col0=col0)
    r = o.conv_tend_batch(ch, nthreads=2)
    w = torch.from_numpy(_budget(o, ch, r))
    dist.all_reduce(w)
    q.put((rank, w.numpy(), r["ptend_q"], r["ideep"], r["lengath"], r["prec"]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank(built):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cam_nor_physics_b200 import soundings as S
    from helpers import get_oracle
    ncols, world = 16 * 13 + 5, 2            # 14 chunks, the last one ragged: ranks get 7 chunks each
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_worker, args=(r, world, port, ncols, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    o, _, _ = get_oracle("libm", 16, 32)
    ch = S.make_chunks(ncols, 32, 16, p_conv=0.5)
    r = o.conv_tend_batch(ch)
    assert np.array_equal(np.concatenate([res[0][2], res[1][2]]), r["ptend_q"])
    assert np.array_equal(np.concatenate([res[0][3], res[1][3]]), r["ideep"])
    assert np.array_equal(np.concatenate([res[0][4], res[1][4]]), r["lengath"])
    assert np.array_equal(np.concatenate([res[0][5], res[1][5]]), r["prec"])
    w1 = _budget(o, ch, r)
    for rk in range(world):
        assert np.allclose(res[rk][1], w1, rtol=1e-13, atol=1e-300)
    assert w1[4] > 0 and w1[5] == ncols


def test_shard_partition_covers_the_grid_in_whole_chunks():
    """bench.shard_of: strong scaling cuts ONE grid into contiguous blocks of whole pcols=16 chunks (chunk c -> rank
    floor(c*world/nchunks)); weak scaling gives every rank its own grid."""
    sys.path.insert(0, ROOT)
    import bench
    for ncols in (55296, 13824, 16 * 13 + 5, 17):
        for world in (1, 2, 3, 4, 8):
            parts = [bench.shard_of(r, world, ncols, "strong") for r in range(world)]
            assert parts[0][0] == 0 and sum(n for _, n in parts) == ncols
            for (c0, n), (c1, _) in zip(parts, parts[1:]):
                assert c0 + n == c1 and c1 % 16 == 0
            assert all(n >= 0 for _, n in parts)
            if ncols % (16 * world) == 0:
                assert len({n for _, n in parts}) == 1
    assert bench.shard_of(3, 8, 55296, "weak") == (3 * 55296, 55296)
