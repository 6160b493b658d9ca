"""N > 1 path on CPU: world_size-2 gloo run of the flat gradient bucket + all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Toy(torch.nn.Module):
    """Stands in for a compressor: main parameters, `quantiles` (excluded from the exchange), a loss balancer."""

    def __init__(self):
        super().__init__()
        self.model = torch.nn.ModuleDict({"net": torch.nn.Linear(6, 3)})
        self.model["net"].quantiles = torch.nn.Parameter(torch.zeros(3, 1, 3))
        self.loss_balancer = torch.nn.Module()
        self.loss_balancer.log_vars = torch.nn.Parameter(torch.zeros(2))
        self.grad_sync = None

    def get_main_parameters(self):
        return [p for n, p in self.model.named_parameters() if not n.endswith(".quantiles")]

    def loss(self, x):
        y = self.model["net"](x)
        return (y ** 2).mean() * torch.exp(-self.loss_balancer.log_vars).sum() + self.loss_balancer.log_vars.sum()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mmnc_b200 as mm

    torch.manual_seed(100 + rank)  # different init per rank: broadcast must fix it
    toy = _Toy()
    dp = mm.DataParallel(toy, n_buckets=3)   # [log_vars, bias | weight]: every all-reduce starts from a grad hook
    assert len(dp.bucket.slices) == 2 and dp.bucket.counts == [2, 1]
    torch.manual_seed(7)
    x_warm, x_all = torch.randn(8, 6), torch.randn(8, 6)
    for x in (x_warm, x_all):                # two steps: the arrival state must re-arm, the bucket must be re-zeroed
        toy.grad_zero()
        toy.loss(x[rank * 4:(rank + 1) * 4]).backward()
        assert dp.bucket.check_views()
        assert all(dp._launched), "every bucket's all-reduce should have been issued during backward"
        toy.grad_sync()
        assert not any(dp._launched) and not dp._work
    res = {"grad": dp.bucket.flat.clone(), "w": toy.model["net"].weight.detach().clone(),
           "offset": mm.ops.noise_source.next(10)[1]}
    if rank == 0:
        # single-process reference on the full batch
        ref = _Toy()
        ref.load_state_dict(toy.state_dict())
        ref.loss(x_all).backward()
        res["ref"] = torch.cat([p.grad.reshape(-1) for p in
                                reversed(ref.get_main_parameters() + list(ref.loss_balancer.parameters()))])
    torch.save(res, os.path.join(out, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_gradient_allreduce_matches_large_batch(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (torch.load(tmp_path / f"r{i}.pt") for i in range(world))
    assert torch.equal(r0["w"], r1["w"]), "parameters were not broadcast from rank 0"
    assert torch.equal(r0["grad"], r1["grad"]), "ranks disagree after the all-reduce"
    assert torch.allclose(r0["grad"], r0["ref"], rtol=1e-5, atol=1e-7), "mean of shard grads != full-batch grad"
    assert (r0["offset"], r1["offset"]) == (0, 10), "per-rank Philox offsets (world-size invariant noise)"


def test_flat_bucket_single_process():
    import mmnc_b200 as mm

    toy = _Toy()
    b = mm.FlatGradBucket(toy.get_main_parameters())
    toy.loss(torch.randn(4, 6)).backward()
    assert b.check_views() and b.flat.abs().sum() > 0
    n = sum(p.numel() for p in toy.get_main_parameters())
    assert b.flat.numel() == n
    for p in toy.get_main_parameters():
        p.grad.zero_()
    assert b.flat.abs().sum() == 0
    # several buckets: contiguous, cover the buffer, every parameter in exactly one
    b3 = mm.FlatGradBucket(toy.get_main_parameters() + [toy.loss_balancer.log_vars], n_buckets=2)
    assert b3.bounds[0][0] == 0 and b3.bounds[-1][1] == b3.flat.numel()
    assert all(b3.bounds[i][1] == b3.bounds[i + 1][0] for i in range(len(b3.bounds) - 1))
    assert sum(b3.counts) == 3 and sorted(set(b3.bucket_of.values())) == list(range(len(b3.bounds)))
    toy.loss(torch.randn(4, 6)).backward()
    b3.zero()
    assert b3.flat.abs().sum() == 0 and b3.check_views()
