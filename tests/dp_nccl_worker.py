"""Worker of tests/test_gpu_dp.py (launched by torch.distributed.run, one rank per GPU, NCCL).

Every rank trains on its own shard with the data-parallel Track-M step (features all-gathered, loss on the rank's
strips with the LSE exchange, adapter gradients all-reduced, fused clip + AdamW; eager launches and the two-graph
replay) and checks it against ONE process running the same model on the concatenated global batch:
loss, summed gradients, updated parameters.  Prints "DP_OK" on rank 0."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import clip_oracle as O  # noqa: E402  (test infrastructure: synthetic inputs, model builder)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from vlm_clip_b200.model_m import CLIPWithAdapters
    from vlm_clip_b200.trainer import CLIPAdapterTrainer

    name = "openai/clip-vit-base-patch32"
    clip = O.build_hf_clip(name, seed=0, vision_layers=2, text_layers=2).to(dev)
    for p in clip.parameters():
        p.requires_grad_(False)
    nl, steps = 6, 5
    N = nl * world

    def global_batch(step):
        pix, ids, mask = O.synthetic_batch(N, seed=40 + step)
        ids[:, 0] = (torch.arange(N) * 13 + step) % 1000
        return {"input_ids": ids.to(dev), "attention_mask": mask.to(dev), "pixel_values": pix.to(dev)}

    def shard(b):
        return {k: v[rank * nl:(rank + 1) * nl].contiguous() for k, v in b.items()}

    def make(seed, dp, graph):
        torch.manual_seed(seed)  # rank-dependent seeds on purpose: the trainer must broadcast rank 0's parameters
        m = CLIPWithAdapters(clip=clip, use_shared_adapters=False).to(dev).train()
        if not dp:
            m.enable_data_parallel(enabled=False)
        tr = CLIPAdapterTrainer(m, [None] * steps, learning_rate=1e-3, output_dir=f"/tmp/vlmclip_dp_{rank}", cuda_graph=graph,
                                graph_warmup_steps=2)
        if not dp:
            m.enable_data_parallel(enabled=False)  # the trainer switches it on when a process group exists
        return m, tr

    results = {}
    trainers = []
    for mode, (dp, graph) in {"single": (False, False), "dp_eager": (True, False), "dp_graph": (True, True)}.items():
        m, tr = make(1 if not dp else 1 + 100 * rank, dp, graph)
        trainers.append(tr)
        if not dp:
            import vlm_clip_b200.trainer as T

            saved = T.allreduce_sum_
            T.allreduce_sum_ = lambda t, group=None: t  # the single-process reference must not average anything
            tr.optimizer  # noqa: B018  (creates the optimiser; its broadcast keeps rank 0's init == seed 1)
        losses = []
        for s in range(steps):
            b = global_batch(s)
            losses.append(tr.training_step(b if not dp else shard(b)).clone())
        if not dp:
            T.allreduce_sum_ = saved
        torch.cuda.synchronize()
        results[mode] = (torch.stack(losses), tr.optimizer.flat.clone(), tr.optimizer.grad.clone())
        if graph:
            assert tr.graph_replays == steps - 2, tr.graph_replays
    ref_l, ref_p, ref_g = results["single"]
    for mode in ("dp_eager", "dp_graph"):
        l, p, g = results[mode]
        dl = (l - ref_l).abs().max().item()
        dg = (g - ref_g).abs().max().item() / (ref_g.abs().max().item() + 1e-12)
        dp_ = (p - ref_p).abs().max().item()
        assert dl < 2e-5, (mode, "loss", dl, l.tolist(), ref_l.tolist())
        assert dg < 2e-4, (mode, "grad", dg)
        assert dp_ < 2e-5, (mode, "params", dp_)
        # replicas stay identical
        chk = p.double().sum().reshape(1)
        both = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(both, chk)
        assert all(torch.equal(both[0], x) for x in both), (mode, both)
        if rank == 0:
            print(f"{mode}: loss diff {dl:.2e}, grad rel diff {dg:.2e}, param diff {dp_:.2e}")
    assert torch.equal(results["dp_eager"][0], results["dp_graph"][0])  # replay == eager, bit for bit
    for tr in trainers:
        tr.release_graphs()  # NCCL cannot tear a communicator down while graphs holding its kernels exist
    dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        print("DP_OK", flush=True)
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
