"""The four model configurations BASELINE.json names, at their REAL sizes (not toy channel counts), against the oracle:

  C1  SingleTaskCompressor            -m 1  mono                                  -l 32  -c 32
  C2  MultiTaskDisjointLatentCompressor -m 3  rgb depth_euclidean normal            -l 128 -c 100   (headline)
  C3  MultiTaskMixedLatentCompressor  -m 2  rgb depth_euclidean normal semantic   -l 192 -c 128
  C4  MultiTaskSharedLatentCompressor -m 4  rgb depth_euclidean normal semantic   -l 192 -c 128

(/root/reference/src/train.py:89-120, /root/reference/README.md:50-57.)  Batch 2 at 256 x 256: the CPU oracle takes
0.1-1 s per pass.  Convolutions run on different devices in the two arms (cuDNN fp32 vs torch CPU), so scalar losses
are held to 5e-4 with the GDN contraction in fp32 and to 5e-3 with the product default ("auto" = single-pass TF32 on
the tensor cores, what the reference's own F.conv2d GDN computes under cuDNN's default).
"""
import pytest
import torch

import mmnc_b200 as mm
from oracle import reference_models as orm

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

ALL4 = ("rgb", "depth_euclidean", "normal", "semantic")
REAL_CONFIGS = {
    "C1": (1, ("mono",), 32, 32),
    "C2": (3, ("rgb", "depth_euclidean", "normal"), 128, 100),
    "C3": (2, ALL4, 192, 128),
    "C4": (4, ALL4, 192, 128),
}


# fused backward variant the GDN(c) @ 128 x 128 layer of each configuration must take (3 = TMA-fed tcgen05 kernel)
WIDE_VARIANT = {100: 3, 128: 4}  # batch 2 here: too few tiles per CTA for the x-prefetch variant (5)


def _pair(name, seed):
    kind, tasks, l, c = REAL_CONFIGS[name]
    torch.manual_seed(seed)
    ours = mm.build_compressor(kind, tasks, l, c, lmbda=1e-2)
    ref = orm.ReferenceCompressor(kind, tasks, l, c, lmbda=1e-2)
    ref.load_state_dict(ours.state_dict())
    return ours, ref, tasks


def _set_precision(model, precision):
    for m in model.modules():
        if isinstance(m, mm.GDN):
            m.precision = precision


@pytest.mark.parametrize("precision", ["fp32", "auto"])
@pytest.mark.parametrize("name", list(REAL_CONFIGS))
def test_real_config_eval_forward_vs_oracle(name, precision):
    ours, ref, tasks = _pair(name, 50)
    ref.eval()
    _set_precision(ours, precision)
    ours.to(DEV).eval()
    B = 2
    batch = mm.synthetic_batch(tasks, B, size=256, seed=21)
    bd = {k: v.to(DEV) for k, v in batch.items()}
    with torch.backends.cudnn.flags(allow_tf32=False), torch.no_grad():
        x_hats, lik = ours(bd)
        loss, logs = ours.rate_distortion_loss(bd, x_hats, lik, "val")
        want, want_logs = ref.rd_loss(batch, "val")
        ref_x_hats, ref_lik = ref(batch)
    M, N = ours.model["compressor"].M, ours.model["compressor"].N
    assert lik["y"].shape == (B, M, 4, 4) == tuple(ref_lik["y"].shape)
    assert lik["z"].shape == (B, N, 1, 1) == tuple(ref_lik["z"].shape)
    for t in tasks:
        assert x_hats[t].shape == ref_x_hats[t].shape
    tol = 5e-4 if precision == "fp32" else 5e-3
    assert abs(loss.item() - want.item()) <= tol * abs(want.item()), (loss.item(), want.item())
    assert set(logs) == set(want_logs)
    for k in want_logs:
        assert abs(float(logs[k]) - float(want_logs[k])) <= 4 * tol * abs(float(want_logs[k])) + 1e-6, k
    if precision == "fp32":
        # the z likelihoods only see convolutions + fp32 GDN upstream: element-wise agreement, floor-aware
        got, exp = lik["z"].cpu().double(), ref_lik["z"].double()
        big = exp > 1e-3
        assert ((got - exp).abs() / exp)[big].max().item() < 2e-2


@pytest.mark.parametrize("name", list(REAL_CONFIGS))
def test_real_config_train_gradients_vs_oracle(name):
    """One training-mode forward + backward with the same injected quantisation noise in both arms: the gradient of
    EVERY parameter tensor against the oracle's autograd (max error relative to the tensor's largest entry)."""
    ours, ref, tasks = _pair(name, 60)
    ref.train()
    _set_precision(ours, "fp32")
    ours.to(DEV).train()
    B = 1 if name == "C3" else 2
    batch = mm.synthetic_batch(tasks, B, size=256, seed=21)
    bd = {k: v.to(DEV) for k, v in batch.items()}
    co, cr = ours.model["compressor"], ref.model["compressor"]
    g = torch.Generator().manual_seed(9)
    nz = torch.rand(B, co.N, 1, 1, generator=g) - 0.5
    ny = torch.rand(B, co.M, 1, 1, generator=g) - 0.5
    eb_f, gc_f, eb_rf, gc_rf = (co.entropy_bottleneck.forward, co.gaussian_conditional.forward,
                                cr.entropy_bottleneck.forward, cr.gaussian_conditional.forward)
    nzd, nyd = nz.to(DEV), ny.to(DEV)
    co.entropy_bottleneck.forward = lambda x, training=None: eb_f(x, training, noise=nzd)
    co.gaussian_conditional.forward = lambda y, s, means=None, training=None: gc_f(y, s, means, training, noise=nyd)
    cr.entropy_bottleneck.forward = lambda x, training=None: eb_rf(x, training, noise=nz)
    cr.gaussian_conditional.forward = lambda y, s, means=None, training=None: gc_rf(y, s, means, training, noise=ny)
    with torch.backends.cudnn.flags(allow_tf32=False):
        x_hats, lik = ours(bd)
        loss, _ = ours.rate_distortion_loss(bd, x_hats, lik, "train")
        loss.backward()
    want, _ = ref.rd_loss(batch, "train")
    want.backward()
    assert abs(loss.item() - want.item()) <= 5e-4 * abs(want.item())
    ref_params = dict(ref.named_parameters())
    checked, worst = 0, (0.0, "")
    for pname, p in ours.named_parameters():
        if p.grad is None or pname.endswith("quantiles"):
            continue
        gr, w = p.grad.cpu(), ref_params[pname].grad
        denom = w.abs().max().item()
        if denom < 1e-12:
            continue
        err = (gr - w).abs().max().item() / denom
        worst = max(worst, (err, pname))
        checked += 1
    assert worst[0] < 2e-2, worst
    assert checked > (40 if name == "C1" else 100)


@pytest.mark.parametrize("name", ["C2", "C3", "C4"])
def test_real_config_training_step_runs_and_uses_tensor_cores(name):
    """The public `training_step` at the real channel counts; and the contraction variant the big layers take."""
    kind, tasks, l, c = REAL_CONFIGS[name]
    torch.manual_seed(70)
    ours = mm.build_compressor(kind, tasks, l, c, lmbda=1e-2).to(DEV).train()
    ours.configure_optimizers(total_steps=4)
    bd = mm.synthetic_batch(tasks, 2, size=256, seed=21, device=DEV)
    l0 = ours.training_step(bd).item()
    l1 = ours.training_step(bd).item()
    assert l0 == l0 and l1 == l1 and l1 < l0 * 1.5
    L = mm._lib.lib()
    x = torch.empty(2, c // 2, 256, 256, device=DEV)
    auto = mm.ops.GDN_PRECISION["auto"]
    assert L.mmnc_gdn_backward_variant(x.data_ptr(), x.data_ptr(), 2, c // 2, 256 * 256, auto) == 3
    x = torch.empty(2, c, 128, 128, device=DEV)
    assert L.mmnc_gdn_backward_variant(x.data_ptr(), x.data_ptr(), 2, c, 128 * 128, auto) == WIDE_VARIANT[c]
    if name == "C3":  # the mixed-latent output heads: IGDN(2 c = 256) at 64 x 64 and 32 x 32 -> the streamed-operand kernels
        x = torch.empty(2, 2 * c, 64, 64, device=DEV)
        assert L.mmnc_gdn_forward_variant(x.data_ptr(), x.data_ptr(), 2, 2 * c, 64 * 64, auto) == 6
        assert L.mmnc_gdn_backward_variant(x.data_ptr(), x.data_ptr(), 2, 2 * c, 64 * 64, auto) == 6
        names = [type(m).__name__ + str(m.beta.numel()) for m in ours.modules() if isinstance(m, mm.GDN)]
        assert "GDN256" in names, "config C3 should contain 256-channel IGDN layers"


def test_real_config_actual_bits_track_estimated_bits():
    """src/check_bpp.ipynb:129-130 as a property, at the C2 latent shapes: the rANS-coded size of a batch of
    latents equals the likelihood estimate to within the coder's per-stream overhead (a 64-bit flush per string)."""
    torch.manual_seed(80)
    kind, tasks, l, c = REAL_CONFIGS["C2"]
    ours = mm.build_compressor(kind, tasks, l, c, lmbda=1e-2)
    ours.update_bottleneck_values()
    ours.to(DEV).eval()
    comp = ours.model["compressor"]
    B = 128
    scales = torch.exp(torch.empty(B, comp.M, 1, 1, device=DEV).uniform_(-2.2, 2.0))
    y = torch.randn(B, comp.M, 1, 1, device=DEV) * scales
    z = torch.randn(B, comp.N, 1, 1, device=DEV) * 3
    with torch.no_grad():
        idx = comp.gaussian_conditional.build_indexes(scales)
        ys = comp.gaussian_conditional.compress(y, idx)
        zs = comp.entropy_bottleneck.compress(z)
        _, ylik = comp.gaussian_conditional(y, scales)
        _, zlik = comp.entropy_bottleneck(z)
    est_bits = float(-torch.log2(ylik).sum() - torch.log2(zlik).sum())
    actual_bits = 8 * (sum(map(len, ys)) + sum(map(len, zs)))
    overhead = 2 * B * 64  # two strings per image, each flushed with 64 bits of state
    assert est_bits * 0.97 <= actual_bits <= est_bits * 1.06 + overhead, (actual_bits, est_bits)
    bpp_actual = actual_bits / (B * 256 * 256 * len(tasks))
    bpp_est = est_bits / (B * 256 * 256 * len(tasks))
    assert f"{bpp_actual:.2f}" == f"{bpp_est:.2f}"  # "equal to the printed digits", as in the notebook
