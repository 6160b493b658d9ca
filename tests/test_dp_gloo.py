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
    dp = mm.DataParallel(toy)
    torch.manual_seed(7)
    x_all = torch.randn(8, 6)
    shard = x_all[rank * 4:(rank + 1) * 4]
    toy.loss(shard).backward()
    assert dp.bucket.check_views()
    toy.grad_sync()
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
