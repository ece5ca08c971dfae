"""CPU: manifest sharding + final gather, world_size 2 and 3 over gloo (the N>1 path of bench.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from taiwan_whisper_b200.shard import gather_token_rows, run_manifest, shard_bounds


def test_shard_bounds_partition():
    for n in (0, 1, 7, 64, 100_000):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_clips, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # fake "transcription": tokens of clip i are [i, i+1, ...] with length 1 + i % 5
        def load_batch(lo, hi):
            return torch.arange(lo, hi, dtype=torch.int16)[:, None].repeat(1, 4)

        def transcribe(pcm):
            ids = pcm[:, 0].to(torch.int32)
            L = 6
            toks = ids[:, None] + torch.arange(L, dtype=torch.int32)[None, :]
            lens = 1 + ids % 5
            toks = torch.where(torch.arange(L)[None, :] < lens[:, None], toks, torch.full_like(toks, -7))
            return toks, lens

        toks, lens = run_manifest(n_clips, 4, load_batch, transcribe, pad_id=-7, rank=rank, world=world)
        ok = toks.shape[0] == n_clips and bool((toks[:, 0] == torch.arange(n_clips, dtype=torch.int32)).all()) and \
            bool((lens == 1 + torch.arange(n_clips, dtype=torch.int32) % 5).all()) and \
            bool(((toks == -7) == (torch.arange(toks.shape[1])[None, :] >= lens[:, None])).all())
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_clips", [(2, 11), (3, 2), (2, 64)])
def test_run_manifest_gloo(world, n_clips):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_clips, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=10) for _ in range(world))
    assert all(res.values()), res


def test_gather_single_process_passthrough():
    t = torch.zeros((3, 5), dtype=torch.int32)
    l = torch.ones((3,), dtype=torch.int32)
    a, b = gather_token_rows(t, l, 0)
    assert a is t and b is l
