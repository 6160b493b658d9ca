"""Data parallelism on real GPUs: needs >= 2 CUDA devices, skipped otherwise.

  gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp.py -m gpu -q

World size 2, NCCL, the real compressor (-m 3, three tasks) at reduced width:
  * rate-path modules are per image: eval outputs of a shard are BIT-equal to the same images inside the full batch;
  * the overlapped bucket all-reduce (parallel.DataParallel) gives every rank the gradient of the 1-GPU large batch
    (same Philox noise: it is keyed by the global element index) within fp32 reduction tolerance;
  * after one optimizer step the replicas still hold identical parameters.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
TASKS = ("rgb", "depth_euclidean", "normal")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _grads(model):
    return {n: p.grad.detach().clone() for n, p in model.named_parameters()
            if p.grad is not None and not n.endswith("quantiles")}


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import mmnc_b200 as mm

    per_rank = 2
    full = mm.synthetic_batch(TASKS, per_rank * world, size=256, seed=21)
    shard = {k: v[rank * per_rank:(rank + 1) * per_rank].to(dev) for k, v in full.items()}

    def build():
        torch.manual_seed(100)  # same init everywhere (DataParallel broadcasts anyway)
        m = mm.build_compressor(3, TASKS, 24, 20, lmbda=1e-2)
        for g in m.modules():
            if isinstance(g, mm.GDN):
                g.precision = "fp32"
        return m.to(dev)

    res = {}
    with torch.backends.cudnn.flags(allow_tf32=False, deterministic=True):
        # ---- (1) eval: rate-path modules on a shard vs the same images inside the full batch (bit-equal)
        model = build().eval()
        c = model.model["compressor"]
        g = torch.Generator().manual_seed(5)
        z_full = torch.randn(per_rank * world, c.N, 1, 1, generator=g) * 3
        s_full = torch.exp(torch.empty(per_rank * world, c.M, 4, 4).uniform_(-2, 3, generator=g))
        y_full = torch.randn(per_rank * world, c.M, 1, 1, generator=g) * 2
        x_full = torch.randn(per_rank * world, 20, 64, 64, generator=g)
        sl = slice(rank * per_rank, (rank + 1) * per_rank)
        gdn = model.model["input_heads"][0][3]
        with torch.no_grad():
            mine = (c.entropy_bottleneck(z_full[sl].to(dev)), c.gaussian_conditional(y_full[sl].to(dev), s_full[sl].to(dev)),
                    gdn(x_full[sl].to(dev)))
            whole = (c.entropy_bottleneck(z_full.to(dev)), c.gaussian_conditional(y_full.to(dev), s_full.to(dev)),
                     gdn(x_full.to(dev)))
        res["eval_bit_equal"] = bool(
            torch.equal(mine[0][0], whole[0][0][sl]) and torch.equal(mine[0][1], whole[0][1][sl]) and
            torch.equal(mine[1][0], whole[1][0][sl]) and torch.equal(mine[1][1], whole[1][1][sl]) and
            torch.equal(mine[2], whole[2][sl]))

        # ---- (2) training gradients: 2-GPU shards + overlapped all-reduce vs the 1-GPU large batch
        model = build().train()
        dp = mm.DataParallel(model, n_buckets=4)
        model.configure_optimizers(total_steps=4)
        torch.manual_seed(7)  # Philox seed (same on every rank; rank r reads its own slice of the noise stream)
        model.grad_zero()
        x_hats, lik = model(shard)
        loss, _ = model.rate_distortion_loss(shard, x_hats, lik, "train")
        # the reference's normalisers use the per-step batch: loss terms are per-image means, so the mean over
        # ranks of the shard gradients is the gradient of the large-batch loss
        loss.backward()
        launched_in_backward = sum(dp._launched)
        model.grad_sync()
        got = _grads(model)
        res["launched_in_backward"], res["buckets"] = launched_in_backward, len(dp.bucket.slices)
        if rank == 0:
            ref = build().train()
            ref.load_state_dict(model.state_dict())
            mm.ops.noise_source.configure(0, 1)
            torch.manual_seed(7)
            fb = {k: v.to(dev) for k, v in full.items()}
            xh, lk = ref(fb)
            l2, _ = ref.rate_distortion_loss(fb, xh, lk, "train")
            l2.backward()
            mm.ops.noise_source.configure(rank, world)
            want = _grads(ref)
            worst = (0.0, "")
            for n, w in want.items():
                d = w.abs().max().item()
                if d > 1e-12:
                    worst = max(worst, ((got[n] - w).abs().max().item() / d, n))
            res["worst_grad_err"] = worst
        # ---- (3) one optimizer step keeps the replicas identical
        model.optimizers()[0].step()
        flat = torch.cat([p.detach().reshape(-1) for p in model.get_main_parameters()])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        res["replicas_equal"] = all(torch.equal(gathered[0], t) for t in gathered[1:])
    torch.save(res, os.path.join(out, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpu_data_parallel_matches_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs (run under gpurun --gpus 2)")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / f"r{i}.pt") for i in range(world)]
    assert all(x["eval_bit_equal"] for x in r), "per-image rate-path outputs depend on the batch they sit in"
    assert all(x["replicas_equal"] for x in r), "replicas diverged after one step"
    assert r[0]["buckets"] >= 2 and r[0]["launched_in_backward"] >= 1, "no all-reduce overlapped the backward pass"
    err, name = r[0]["worst_grad_err"]
    assert err < 5e-5, f"all-reduced gradient differs from the large-batch gradient: {err:.2e} of max at {name}"
