"""Host-side data-parallel logic on CPU with the gloo backend, world_size 2:
bucketed asynchronous gradient all-reduce == mean of the per-rank gradients, the replicas stay bit-identical
after an optimizer step, inference shards cover the utterances exactly once, and the dropout stream of
rank r is the slice [r*B, (r+1)*B) of the global batch's stream."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from genvox_b200.training import allreduce_gradients, bucketize, shard_rows
from oracle import philox


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                                  # identical init on every rank
        model = torch.nn.Sequential(torch.nn.Linear(37, 64), torch.nn.Tanh(), torch.nn.Linear(64, 5))
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-6)
        g = torch.Generator().manual_seed(100 + rank)         # different shard per rank
        x, y = torch.randn(8, 37, generator=g), torch.randn(8, 5, generator=g)
        loss = torch.nn.functional.mse_loss(model(x), y)
        loss.backward()
        local = [p.grad.clone() for p in model.parameters()]
        nb = allreduce_gradients(list(model.parameters()), bucket_mb=0.005)     # ~5 KB buckets -> several buckets
        reduced = [p.grad.numpy().copy() for p in model.parameters()]
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        out.put((rank, nb, [t.numpy() for t in local], reduced, [p.detach().numpy().copy() for p in model.parameters()]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_bucketed_allreduce_averages_gradients_and_keeps_replicas_identical():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, nb0, loc0, red0, par0), (_, nb1, loc1, red1, par1) = res
    assert nb0 == nb1 and nb0 >= 2                                      # really bucketed
    for a, b, r0, r1 in zip(loc0, loc1, red0, red1):
        np.testing.assert_allclose(r0, (a + b) / 2, rtol=1e-6, atol=1e-7)   # mean over ranks
        assert np.array_equal(r0, r1)
    for p0, p1 in zip(par0, par1):
        assert np.array_equal(p0, p1)                                   # replicas stay in lock-step


def test_bucketize_keeps_order_and_bounds_payload():
    ts = [torch.zeros(n) for n in (10, 2000, 30, 5000, 1, 1)]
    buckets = bucketize(ts, 4096 * 4)
    assert [t.numel() for b in buckets for t in b] == [10, 2000, 30, 5000, 1, 1]
    assert all(sum(t.numel() for t in b) * 4 <= 4096 * 4 or len(b) == 1 for b in buckets)
    assert allreduce_gradients([]) == 0                                # no process group: no-op


def test_inference_shards_cover_every_utterance_once():
    for n, world in ((4096, 8), (10, 3), (7, 8), (64, 1)):
        spans = [shard_rows(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_rank_dropout_stream_is_a_slice_of_the_global_batch_stream():
    full = philox.keep_mask(9, philox.SITE_DEC, 5, 128, 1024, 0.1)
    for rank in range(2):
        part = philox.keep_mask(9, philox.SITE_DEC, 5, 64, 1024, 0.1, row_offset=rank * 64)
        assert np.array_equal(part, full[rank * 64:(rank + 1) * 64])
