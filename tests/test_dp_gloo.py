"""World-size-2 `gloo` test of the data-parallel protocol on the CPU: all-gather of the local
features, global loss on every rank, gradients for local rows only, SUM all-reduce of parameter gradients —
must reproduce the single-process loss and gradients of the concatenated batch.  The arithmetic here is the
oracle's (the CUDA kernel's local-row gradient is checked against the same identity on the GPU in
tests/test_gpu_kernels.py::test_clip_loss_data_parallel_rows)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import clip_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vlm_clip_b200.dist import allreduce_sum_, gather_features

    torch.manual_seed(0)
    N, P, D = 8, 32, 16
    nl = N // world
    x_t, x_i = torch.randn(N, D), torch.randn(N, D)  # identical on every rank (same seed): the global batch
    W = torch.nn.Parameter(torch.randn(P, D) * 0.3)  # a replicated trainable "adapter"
    scale = torch.tensor(2.0)
    # local forward on this rank's shard
    sl = slice(rank * nl, (rank + 1) * nl)
    t_loc, i_loc = x_t[sl] @ W.t(), torch.tanh(x_i[sl] @ W.t())
    t_all, i_all, row0 = gather_features(t_loc, i_loc)
    assert row0 == rank * nl and t_all.shape == (N, P)
    loss = O.contrastive_loss_local_rows(t_all, i_all, t_loc, i_loc, row0, scale)
    loss.backward()
    flat = W.grad.reshape(-1).clone()
    allreduce_sum_(flat)
    # single-process reference on the concatenated batch
    W2 = torch.nn.Parameter(W.detach().clone())
    ref = O.contrastive_loss(x_t @ W2.t(), torch.tanh(x_i @ W2.t()), scale)["loss"]
    ref.backward()
    q.put((rank, abs(loss.item() - ref.item()), (flat - W2.grad.reshape(-1)).abs().max().item(),
           W2.grad.abs().max().item()))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_dp_global_loss_and_summed_grads_match_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, dl, dg, gmax in res:
        assert dl < 1e-6, (rank, dl)
        assert dg < 1e-6 * max(1.0, gmax) + 1e-7, (rank, dg, gmax)


def _bcast_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vlm_clip_b200 import ops

    torch.manual_seed(100 + rank)  # replicas built with DIFFERENT RNG state (nothing synchronised them)
    params = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(7))]
    opt = ops.FusedAdamW(params, lr=1e-3)
    opt.exp_avg.fill_(float(rank + 1))
    opt.step_t.fill_(rank + 3)
    gen0 = ops.param_generation()
    opt.broadcast_from(0)
    # plain lists: tensors sent through a Queue are shared-memory handles that die with this process
    q.put((rank, opt.flat.tolist(), opt.exp_avg.tolist(), int(opt.step_t.item()), params[0].detach().reshape(-1).tolist(),
           ops.param_generation() - gen0))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_dp_replicas_start_from_rank0_parameters():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bcast_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=100) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    (_, f0, m0, s0, w0, g0), (_, f1, m1, s1, w1, g1) = res
    assert f0 == f1 and m0 == m1 and s0 == s1 == 3
    assert w0 == w1  # the nn.Parameters are views of the arena: they follow it
    torch.manual_seed(100)
    assert w0 == torch.randn(5, 3).reshape(-1).tolist()  # ... and it is rank 0's initialisation that won
    assert g0 == 1 and g1 == 1  # cached weight packs are invalidated


def _bucket_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vlm_clip_b200 import ops
    from vlm_clip_b200.dist import BucketedGradAllReduce

    torch.manual_seed(0)
    # arena order: [emb] [layer 0: 3 tensors] [layer 1: 3 tensors] [head, a scalar]  (odd sizes exercise the 16-byte padding)
    emb = torch.nn.Parameter(torch.randn(7, 3))
    layers = [[torch.nn.Parameter(torch.randn(5, 5)), torch.nn.Parameter(torch.randn(5)), torch.nn.Parameter(torch.randn(2, 5))]
              for _ in range(2)]
    head = torch.nn.Parameter(torch.randn(()))
    params = [emb] + layers[0] + layers[1] + [head]
    opt = ops.FusedAdamW(params, lr=1e-3)
    opt.zero_grad()
    b = BucketedGradAllReduce(opt)
    g = torch.Generator().manual_seed(10 + rank)  # every rank has its own gradients
    mine = {id(p): torch.randn(p.shape, generator=g) for p in params}
    for layer in reversed(layers):                 # the backward hands the layers over last to first ...
        b.sink(layer, [mine[id(p)] for p in layer])
    emb.grad.copy_(mine[id(emb)])                  # ... and autograd accumulates the rest on its own
    head.grad.copy_(mine[id(head)])
    n_calls = b.finish()
    # plain lists: tensors sent through a Queue are shared-memory handles that die with this process
    q.put((rank, n_calls, [p.grad.reshape(-1).tolist() for p in params], [mine[id(p)].reshape(-1).tolist() for p in params]))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_bucketed_gradient_allreduce_covers_the_arena_once():
    """dist.BucketedGradAllReduce (config 5 under data parallelism): layer buckets started from inside the backward plus
    the remainder at the end must give every parameter the SUM of the ranks' gradients, exactly once."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=100) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    (_, n0, got0, mine0), (_, n1, got1, mine1) = res
    assert n0 == n1 == 4  # two layer buckets + the ranges before the layers and after them
    for a, b_, m0, m1 in zip(got0, got1, mine0, mine1):
        assert a == b_
        assert torch.allclose(torch.tensor(a), torch.tensor(m0) + torch.tensor(m1), atol=1e-6)


def test_gather_is_identity_without_process_group():
    from vlm_clip_b200.dist import allreduce_sum_, gather_features, world

    assert world() == (1, 0)
    t, i = torch.randn(4, 8), torch.randn(4, 8)
    ta, ia, r0 = gather_features(t, i)
    assert r0 == 0 and torch.equal(ta, t) and torch.equal(ia, i)
    g = torch.ones(3)
    assert allreduce_sum_(g) is g
