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


# ---- shard-boundary halo: windows of the global concatenation under clip sharding (SURVEY.md §8(e) nuance) ----
def test_shard_window_plan_partitions_the_global_windows():
    sys.path.insert(0, ROOT)
    from seld_b200.dataset import shard_window_plan
    W, H = 250, 50
    for frames in ([327], [202, 125], [3000, 3000, 3000], [30, 40, 5, 700], [0, 260, 0, 10], [49, 1, 50, 100]):
        total = sum(frames)
        want = list(range(0, total, H))  # dataset.py:274-314: start += hop while start < total
        got = []
        for r in range(len(frames)):
            off, starts, halo, first = shard_window_plan(frames, r, W, H)
            assert off == sum(frames[:r])
            g = [off + s for s in starts]
            assert all(off <= x < off + frames[r] for x in g)
            if g:
                assert first == g[0] // H
                assert halo == max(0, min(g[-1] + W, total) - (off + frames[r]))
            got += g
        assert got == want


def _halo_worker(rank, world, port, frames, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from seld_b200.dataset import exchange_shard_halo
    W, H, row_len = 250, 50, 12
    off = sum(frames[:rank])
    T = frames[rank]
    feat = (torch.arange(off, off + T, dtype=torch.float32).unsqueeze(1) * 100 + torch.arange(row_len, dtype=torch.float32))
    # one event every 7 global frames, 5 rows long, class = frame % 13, cell = frame % 648 (rows local to the rank)
    g0 = np.arange(-(-off // 7) * 7, off + T, 7)
    ev = np.stack([g0 - off, np.minimum(g0 - off + 5, T), g0 % 13, g0 % 648], 1).astype(np.int32) if len(g0) else np.zeros((0, 4), np.int32)
    h = min(T, W)
    sel = ev[:, 0] < h
    head_ev = ev[sel].copy()
    head_ev[:, 1] = np.minimum(head_ev[:, 1], h)
    plan, halo_feat, halo_ev, halo_ce = exchange_shard_halo(feat[:h], head_ev, None, T, W, H)
    q.put((rank, plan, halo_feat.numpy(), halo_ev))
    dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("frames", [[327, 260], [120, 90], [600, 0]])
def test_halo_exchange_two_ranks_gloo(frames):
    """Two ranks, gloo: every rank's windows + halo reproduce the rows (and the events) the unsharded concatenation has
    at those positions, including a halo that spans a whole short shard and an empty rank."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_halo_worker, args=(r, world, port, frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=150) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    W, H, row_len, total = 250, 50, 12, sum(frames)
    full = (np.arange(total, dtype=np.float32)[:, None] * 100 + np.arange(row_len, dtype=np.float32))
    n_windows = 0
    for rank, (off, starts, halo, first), halo_feat, halo_ev in res:
        T = frames[rank]
        local = np.concatenate([full[off:off + T], halo_feat]) if T + halo else np.zeros((0, row_len), np.float32)
        assert halo_feat.shape == (halo, row_len)
        for s in starts:
            n = min(W, total - (off + s))
            assert s + n <= T + halo                      # every real frame of the window is available locally
            assert np.array_equal(local[s:s + n], full[off + s:off + s + n])
        # halo events: exactly the global events that touch the halo rows, in local coordinates
        want = [(g - off, min(g + 5, off + T + halo) - off, g % 13, g % 648)
                for g in range(0, total, 7) if off + T <= g < off + T + halo]
        got = sorted(map(tuple, halo_ev.tolist()))
        assert got == sorted(want)
        n_windows += len(starts)
    assert n_windows == -(-total // H)
