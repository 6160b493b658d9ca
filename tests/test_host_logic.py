"""Host-side mirror of the reference interface: state-dict compatibility with the oracle (= CompressAI's names),
bit-exact CDF tables, error behaviour, wrapper topology and channel groups.  No GPU needed."""
import json
import os

import pytest
import torch

import mmnc_b200 as mm
from oracle import compressai_ref as R
from oracle import reference_models as orm

CONFIGS = [  # (kind, tasks, l, c): BASELINE.json configs C1-C4 at reduced width
    (1, ("mono",), 8, 8),
    (2, ("rgb", "depth_euclidean", "normal", "semantic"), 12, 8),
    (3, ("rgb", "depth_euclidean", "normal"), 14, 12),
    (4, ("rgb", "depth_euclidean", "normal", "semantic"), 13, 8),
]


def test_state_dict_names_match_compressai():
    ours, ref = mm.ScaleHyperprior(8, 12).state_dict(), R.ScaleHyperprior(8, 12).state_dict()
    assert sorted(ours) == sorted(ref)
    for k in ours:
        assert ours[k].shape == ref[k].shape and ours[k].dtype == ref[k].dtype, k
    # SURVEY.md A.9 spot checks
    for k in ("entropy_bottleneck._matrix0", "entropy_bottleneck._factor3", "entropy_bottleneck.quantiles",
              "entropy_bottleneck.target", "entropy_bottleneck.likelihood_lower_bound.bound",
              "gaussian_conditional.scale_bound", "gaussian_conditional.lower_bound_scale.bound",
              "g_a.1.beta", "g_a.1.gamma", "g_a.1.beta_reparam.pedestal", "g_a.1.gamma_reparam.lower_bound.bound",
              "h_s.4.weight"):
        assert k in ours, k


@pytest.mark.parametrize("kind,tasks,l,c", CONFIGS)
def test_wrapper_topology_matches_oracle(kind, tasks, l, c):
    ours = mm.build_compressor(kind, tasks, l, c, lmbda=1e-2)
    ref = orm.ReferenceCompressor(kind, tasks, l, c, lmbda=1e-2)
    a, b = ours.state_dict(), ref.state_dict()
    assert sorted(a) == sorted(b)
    assert all(a[k].shape == b[k].shape for k in a)
    ref.load_state_dict(a)
    assert ours.model["compressor"].M == ref.M
    aux = [n for n, _ in ours.model.named_parameters() if n.endswith(".quantiles")]
    assert aux == ["compressor.entropy_bottleneck.quantiles"]
    assert len(ours.get_auxiliary_parameters()) == 1


def test_reference_channel_groups():
    # C2: l=128, T=3 -> groups of 42, channels 126-127 orphaned
    m = mm.build_compressor(3, ("rgb", "depth_euclidean", "normal"), 128, 6)
    chan, norm, names = m._rate_groups()
    assert len(chan) == 128 and chan[:42] == [0] * 42 and chan[84:126] == [2] * 42 and chan[126:] == [-1, -1]
    assert names == ["rgb", "depth_euclidean", "normal"] and norm == [0, 1, 2]
    # C4: l=192, T=4 -> 190 channels, 5 groups of 38, last one shared
    m = mm.build_compressor(4, ("rgb", "depth_euclidean", "normal", "semantic"), 192, 8)
    chan, norm, names = m._rate_groups()
    assert m.model["compressor"].M == 190 and len(chan) == 190 and chan[-38:] == [4] * 38 and chan[:38] == [0] * 38
    assert names[-1] == "shared" and norm[-1] == 0
    # Mixed: one group
    m = mm.build_compressor(2, ("rgb", "normal"), 16, 8)
    assert m._rate_groups()[0] == [0] * 16


def test_tables_bit_exact_with_oracle():
    ours, ref = mm.GaussianConditional(None), R.GaussianConditional(None)
    assert ours.update_scale_table(mm.get_scale_table()) and ref.update_scale_table(R.get_scale_table())
    assert not ours.update_scale_table(mm.get_scale_table())  # no-op the second time (A.4)
    for k in ("_quantized_cdf", "_cdf_length", "_offset", "scale_table"):
        assert torch.equal(getattr(ours, k), getattr(ref, k)), k
    from helpers import perturb_eb_

    eo, er = mm.EntropyBottleneck(9), R.EntropyBottleneck(9)
    perturb_eb_(er)
    eo.load_state_dict(er.state_dict())
    assert eo.update() and er.update() and not eo.update()
    for k in ("_quantized_cdf", "_cdf_length", "_offset"):
        assert torch.equal(getattr(eo, k), getattr(er, k)), k
    assert eo.update(force=True)


def test_error_behaviour_matches_compressai():
    gc = mm.GaussianConditional(None)
    with pytest.raises(ValueError, match="Uninitialized"):
        gc.compress(torch.zeros(1, 2, 2, 2), torch.zeros(1, 2, 2, 2, dtype=torch.int32))
    gc.update_scale_table(mm.get_scale_table())
    with pytest.raises(ValueError, match="same size"):
        gc.compress(torch.zeros(2, 4, 1, 1), torch.zeros(2, 4, 4, 4, dtype=torch.int32))
    with pytest.raises(ValueError, match="at least 2"):
        gc.compress(torch.zeros(4), torch.zeros(4, dtype=torch.int32))
    with pytest.raises(ValueError, match="Invalid quantization mode"):
        gc.quantize(torch.zeros(1, 1), "nearest")
    with pytest.raises(ValueError):
        mm.GaussianConditional([3.0, 1.0])
    with pytest.raises(ValueError):
        mm.GaussianConditional(None, scale_bound=-1)
    with pytest.raises(ValueError, match="strings"):
        gc.decompress(b"xx", torch.zeros(1, 2, 2, 2, dtype=torch.int32))
    eb = mm.EntropyBottleneck(3)
    with pytest.raises(ValueError, match="Uninitialized"):
        eb.compress(torch.zeros(1, 3, 2, 2))
    with pytest.raises(NotImplementedError):
        mm.EntropyBottleneck(3, filters=(3, 3))


def test_checkpoint_roundtrip_after_update():
    m = mm.ScaleHyperprior(4, 6)
    m.update()
    sd = m.state_dict()
    assert sd["gaussian_conditional._quantized_cdf"].shape == (64, 3133)
    fresh = mm.ScaleHyperprior(4, 6)
    fresh.load_state_dict(sd)
    assert torch.equal(fresh.entropy_bottleneck._quantized_cdf, m.entropy_bottleneck._quantized_cdf)
    # and a training checkpoint (empty CDF buffers) loads too
    mm.ScaleHyperprior(4, 6).load_state_dict(mm.ScaleHyperprior(4, 6).state_dict())


def test_oracle_wrapper_losses_golden(golden_dir):
    want = json.load(open(os.path.join(golden_dir, "wrapper_losses.json")))
    for kind, tasks, l, c in ((1, ("mono",), 8, 8), (3, ("rgb", "depth_euclidean", "normal"), 14, 12)):
        torch.manual_seed(100 + kind)
        m = orm.ReferenceCompressor(kind, tasks, l, c, lmbda=1e-2).eval()
        with torch.no_grad():
            loss, logs = m.rd_loss(orm.synthetic_batch(tasks, 1, size=256, seed=21), "val")
        assert abs(float(loss) - want[str(kind)]["loss"]) <= 2e-5 * abs(want[str(kind)]["loss"])
        assert set(logs) | {"loss"} == set(want[str(kind)])


def test_synthetic_batches_agree():
    tasks = ("rgb", "depth_euclidean", "normal", "semantic")
    a, b = mm.synthetic_batch(tasks, 2, size=32), orm.synthetic_batch(tasks, 2, size=32)
    assert all(torch.equal(a[t], b[t]) for t in tasks)
    assert a["semantic"].max() <= 16 and a["depth_euclidean"].max() <= 4.1


def test_compressor_is_copyable_and_requires_an_optimizer_horizon():
    """runtime handles (CUDA streams, group tables) live outside the nn.Module, so deepcopy / pickling of
    the whole module keep working after use; `configure_optimizers` needs the run's step count."""
    import copy
    import pickle

    m = mm.build_compressor(3, ("rgb", "depth_euclidean", "normal"), 14, 12)
    m._cache[("probe", "cpu")] = (object(),)      # what a forward would leave behind
    twin = copy.deepcopy(m)
    assert twin._cache == {} and ("probe", "cpu") in m._cache
    assert sorted(pickle.loads(pickle.dumps(m)).state_dict()) == sorted(m.state_dict())
    assert "_cache" not in m.__dict__ and not any("stream" in k.lower() for k in m.__dict__)
    with pytest.raises(RuntimeError, match="configure_optimizers"):
        m.optimizers()
    with pytest.raises(TypeError):
        m.configure_optimizers()
    with pytest.raises(ValueError):
        m.configure_optimizers(total_steps=0)
    m.configure_optimizers(total_steps=7)
    assert m.lr_schedulers().T_max == 7
    m.concurrent_heads = False
    assert mm.build_compressor(1, ("mono",), 8, 8).concurrent_heads  # per-instance switch
