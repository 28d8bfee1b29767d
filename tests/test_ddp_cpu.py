"""world_size-2 gloo test of the data-parallel plumbing the build owns (train.ParamArena / GradReducer)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from svd_xtend_b200.train import GradReducer, ParamArena
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(37, 53), torch.nn.Linear(53, 11), torch.nn.Linear(11, 200))
        net[1].requires_grad_(False)               # frozen parameters stay outside the arena
        before = [p.detach().clone() for p in net.parameters()]
        arena = ParamArena(net)
        assert all(torch.equal(a, b) for a, b in zip(before, net.parameters()))   # values preserved, now views
        assert len(arena.params) == 4 and all(p.data_ptr() >= arena.data.data_ptr() for p in arena.params)
        red = GradReducer(arena, bucket_mb=0.004)    # tiny buckets => several of them
        assert len(red.buckets) >= 2
        for step in range(2):
            arena.zero_grad()
            for i, p in enumerate(arena.params):
                arena.grad_views[p].fill_(float((rank + 1) * (i + 1) + step))
            # gradients become ready in reverse order, like a backward pass
            for p in reversed(arena.params):
                red.on_grads_ready([p])
            red.finish()
            for i, p in enumerate(arena.params):
                expect = sum((r + 1) * (i + 1) + step for r in range(world)) / world
                assert torch.allclose(arena.grad_views[p], torch.full_like(p, expect)), (rank, i, step)
        arena.attach_grads()
        assert all(p.grad is arena.grad_views[p] for p in arena.params)
        # ---- sharded optimizer plumbing (ZeRO-1): equal aligned shards, reduce-scatter leaves the SUM of this rank's slice,
        # the in-place all-gather assembles every rank's slice (the AdamW kernel itself is covered on the GPU)
        from svd_xtend_b200.train import ShardedAdamW
        net2 = torch.nn.Sequential(torch.nn.Linear(37, 53), torch.nn.Linear(53, 200))
        arena2 = ParamArena(net2, pad_to=world * 64)
        assert arena2.numel % (world * 64) == 0
        opt = ShardedAdamW(arena2, lr=1e-3)
        assert (opt.hi - opt.lo) * world == arena2.numel and opt.lo == rank * (arena2.numel // world) and opt.m.numel() == opt.hi - opt.lo
        arena2.grad.copy_(torch.arange(arena2.numel, dtype=torch.float32) * (rank + 1))
        shard = opt.reduce_scatter_grads()
        expect = torch.arange(arena2.numel, dtype=torch.float32)[opt.lo:opt.hi] * sum(r + 1 for r in range(world))
        assert torch.equal(shard, expect)
        flat = torch.full((arena2.numel,), -1.0)
        flat[opt.lo:opt.hi] = float(rank + 10)
        opt.all_gather_(flat)
        n = arena2.numel // world
        assert all(torch.all(flat[r * n:(r + 1) * n] == float(r + 10)) for r in range(world))
        try:
            ShardedAdamW(ParamArena(torch.nn.Linear(3, 5, bias=False)), lr=1e-3)      # a 64-float arena: not divisible by world * 64
            raise AssertionError("unpadded arena accepted")
        except ValueError:
            pass
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_grad_reducer_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
