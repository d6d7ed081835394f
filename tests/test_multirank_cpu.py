"""World-size-2 test of the path's only collective on CPU (gloo): clips are sharded by rank with no data-path
exchange, each rank accumulates its scaler partials, one all_reduce of [sum | sum of squares | count] gives every rank
the statistics of the whole corpus.  Also the contiguous-block shard rule bench.py / the dataset use."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, feats, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from seld_b200.scaler import FeatureScaler
    from seld_b200.dataset import shard_clips
    lo, hi = shard_clips(len(feats), rank, world)
    sc = FeatureScaler(feats[0].shape[1], device="cpu")
    for f in feats[lo:hi]:  # the partials the feature kernel would have accumulated for this rank's clips
        x = f.double()
        sc.merge(torch.cat([x.sum(0), (x * x).sum(0)]), x.shape[0])
    sc.sync()
    mean, std = sc.finalize()
    q.put((rank, lo, hi, mean.numpy(), std.numpy(), float(sc.buf[-1])))
    dist.destroy_process_group()


def test_shard_clips_contiguous_balanced():
    sys.path.insert(0, ROOT)
    from seld_b200.dataset import shard_clips
    for n in (0, 1, 7, 8, 600):
        for world in (1, 2, 4, 8):
            parts = [shard_clips(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def test_shard_files_uses_rank_env(monkeypatch):
    sys.path.insert(0, ROOT)
    from seld_b200.dataset import shard_files
    a, m = [f"a{i}.wav" for i in range(7)], [f"a{i}.csv" for i in range(7)]
    monkeypatch.setenv("RANK", "1")
    monkeypatch.setenv("WORLD_SIZE", "3")
    assert shard_files(a, m) == (a[3:5], m[3:5])
    got = [shard_files(a, m, r, 3) for r in range(3)]
    assert sum((g[0] for g in got), []) == a and sum((g[1] for g in got), []) == m


@pytest.mark.timeout(120)
def test_scaler_allreduce_two_ranks_gloo():
    g = torch.Generator().manual_seed(0)
    feats = [torch.randn(50 + 7 * i, 448, generator=g) * (1 + i) - 30.0 for i in range(5)]
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, feats, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    allx = torch.cat(feats).double().numpy()
    for rank, lo, hi, mean, std, count in res:
        assert count == allx.shape[0]
        assert np.allclose(mean, allx.mean(0), rtol=1e-12, atol=1e-12)
        assert np.allclose(std, allx.std(0), rtol=1e-9, atol=1e-12)
    assert sorted((lo, hi) for _, lo, hi, *_ in res) == [(0, 3), (3, 5)]
