"""world_size-2 gloo test of the N>1 path's host logic on CPU (no GPU, no NCCL).

Each rank takes its time shard from firpfbch2_time_shards, primes a fresh channelizer with the
halo (here: the CPU oracle stands in for the device kernel -- tests may use it as the checker),
runs its shard, and rank 0 gathers and compares with a single-object pass.  This is exactly
what bench.py does per rank with the CUDA path, including the barrier + max-over-ranks timing.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, M, m, n_frames, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import stimulus
    from oracle import pyoracle as po
    from yagi_b200.sharding import firpfbch2_time_shards

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = firpfbch2_time_shards(n_frames, M, m, world)[rank]
    halo = stimulus.noise_plus_tones(sh.halo_begin, sh.halo_len, M)
    x = stimulus.noise_plus_tones(sh.sample_begin, sh.n_samples, M)
    ch = po.FirPfbCh2.new_kaiser(po.ANALYZER, M, m, 60.0)
    # prime with the halo: (4m-1) frames is odd, so push one extra leading zero-frame to keep parity even
    ch.execute_block(np.concatenate([np.zeros(M // 2, dtype=np.complex64), halo]))
    dist.barrier()
    y = ch.execute_block(x)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                 # the max-over-ranks reduction bench.py uses
    parts = [None] * world
    dist.all_gather_object(parts, (sh.frame_begin, y))
    # the optional output gather (yagi_b200.gather) over the same process group
    from yagi_b200.gather import all_gather_frames, channel_major
    shards = firpfbch2_time_shards(n_frames, M, m, world)
    g = all_gather_frames(torch.from_numpy(y), [s.n_frames for s in shards], M)
    assert tuple(g.shape) == (n_frames, M)
    assert tuple(channel_major(g).shape) == (M, n_frames)
    if rank == 0:
        q.put((t.item(), parts, g.numpy().reshape(-1)))
    dist.destroy_process_group()


@pytest.mark.parametrize("M,m,n_frames", [(16, 5, 402), (64, 3, 251)])
def test_time_sharded_analysis_equals_single_pass(M, m, n_frames):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import stimulus
    from oracle import pyoracle as po

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, M, m, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    tmax, parts, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0
    parts.sort(key=lambda t: t[0])
    y = np.concatenate([p[1] for p in parts])
    x = stimulus.noise_plus_tones(0, n_frames * M // 2, M)
    whole = po.FirPfbCh2.new_kaiser(po.ANALYZER, M, m, 60.0).execute_block(x)
    np.testing.assert_allclose(y, whole, atol=2e-6)
    np.testing.assert_array_equal(gathered, y)
