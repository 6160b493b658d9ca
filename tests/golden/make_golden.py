"""Generates the committed fixtures under tests/golden/ FROM THE ORACLE (oracle/), in this container.

The reference has no golden vectors for this path and CompressAI 1.2.4 is absent, so these
fixtures pin the oracle's behaviour (and thereby every later change to oracle or kernels), not CompressAI's:
parity stays "unpinned" in the sense of DESIGN.md.  Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import compressai_ref as R  # noqa: E402
from oracle import native  # noqa: E402
from oracle import reference_models as orm  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def perturbed_eb(C, seed):
    g = torch.Generator().manual_seed(seed)
    eb = R.EntropyBottleneck(C)
    with torch.no_grad():
        for i in range(5):
            m = getattr(eb, f"_matrix{i}")
            m.add_(torch.randn(m.shape, generator=g) * 0.3)
            b = getattr(eb, f"_bias{i}")
            b.copy_(torch.rand(b.shape, generator=g) - 0.5)
            if i < 4:
                f = getattr(eb, f"_factor{i}")
                f.copy_(torch.rand(f.shape, generator=g) * 2 - 1)
        q = torch.sort(torch.randn(C, 1, 3, generator=g) * 5, dim=2).values
        eb.quantiles.copy_(q)
    return eb


def main():
    torch.manual_seed(1234)
    g = torch.Generator().manual_seed(1234)
    fx = {}

    # ---- known answers -----------------------------------------------------------
    gc = R.GaussianConditional(None)
    gc.update_scale_table(R.get_scale_table())
    eb0 = R.EntropyBottleneck(4)
    eb0.update()
    known = {
        "gc_table_shape": list(gc._quantized_cdf.shape),
        "gc_ragged_entries": int(gc._cdf_length.sum()),
        "gc_pmf_center_first6": (-gc._offset[:6]).tolist(),
        "gc_pmf_center_last3": (-gc._offset[-3:]).tolist(),
        "gc_table_crc": int(np.bitwise_xor.reduce((gc._quantized_cdf.numpy().astype(np.uint64).ravel()
                                                  * (np.arange(gc._quantized_cdf.numel(), dtype=np.uint64) | 1)))),
        "eb_init_cdf_length": eb0._cdf_length.tolist(),
        "eb_init_offset": eb0._offset.tolist(),
        "eb_target": float(eb0.target[2]),
        "eb_matrix_fill_fan3": float(eb0._matrix0[0, 0, 0]),
        "eb_matrix_fill_last": float(eb0._matrix4[0, 0, 0]),
    }
    json.dump(known, open(os.path.join(OUT, "known_answers.json"), "w"), indent=1)

    # ---- pmf_to_quantized_cdf ----------------------------------------------------------------------------
    pmfs, cdfs = [], []
    for n in (2, 3, 7, 23, 64, 301):
        for kind in ("flat", "peaked", "sparse"):
            if kind == "flat":
                p = torch.rand(n, generator=g)
            elif kind == "peaked":
                p = torch.exp(-0.5 * ((torch.arange(n) - n / 2) / max(1.0, n / 12)) ** 2)
            else:
                p = torch.rand(n, generator=g) ** 12
            p = (p / p.sum()).float()
            pmfs.append(p.numpy())
            cdfs.append(np.array(native.pmf_to_quantized_cdf(p.tolist(), 16), dtype=np.int32))
    fx["pmf"] = np.array(pmfs, dtype=object)
    fx["pmf_cdf"] = np.array(cdfs, dtype=object)

    # ---- rANS streams (GC tables, with escapes) ----------------------------------------------------------
    cdf, ln, off = gc._quantized_cdf.numpy(), gc._cdf_length.numpy(), gc._offset.numpy()
    streams = []
    for n in (2, 5, 33, 128, 300, 1024):
        idx = torch.randint(0, 64, (n,), generator=g).int().numpy()
        sc = R.get_scale_table().numpy()[idx]
        sym = np.round(torch.randn(n, generator=g).numpy() * sc).astype(np.int32)
        if n >= 5:
            sym[1], sym[n // 2], sym[-1] = 70000, -70000, int(-off[idx[-1]] + 1)  # escapes incl. boundary
        b = native.encode_with_indexes_np(sym, idx, cdf, ln, off)
        assert (native.decode_with_indexes_np(b, idx, cdf, ln, off) == sym).all()
        streams.append((sym, idx, np.frombuffer(b, dtype=np.uint8)))
    fx["rans_sym"] = np.array([s[0] for s in streams], dtype=object)
    fx["rans_idx"] = np.array([s[1] for s in streams], dtype=object)
    fx["rans_bytes"] = np.array([s[2] for s in streams], dtype=object)

    # ---- EntropyBottleneck (perturbed parameters so the tanh gates are live) -----------------------------
    eb = perturbed_eb(6, 7)
    eb.update()
    z = torch.randn(5, 6, 3, 2, generator=g) * 4
    z[0, 0, 0, 0], z[1, 2, 1, 1] = 60.0, -45.0
    noise = torch.rand(z.shape, generator=g) - 0.5
    eb.eval()
    zh, zl = eb(z)
    eb.train()
    zh_t, zl_t = eb(z, noise=noise)
    zs = eb.compress(z)
    fx.update(eb_state={k: v.numpy() for k, v in eb.state_dict().items()}, eb_z=z.numpy(), eb_noise=noise.numpy(),
              eb_eval_out=zh.detach().numpy(), eb_eval_lik=zl.detach().numpy(), eb_train_out=zh_t.detach().numpy(),
              eb_train_lik=zl_t.detach().numpy(), eb_aux_loss=eb.loss().detach().numpy(),
              eb_strings=np.array([np.frombuffer(s, dtype=np.uint8) for s in zs], dtype=object))

    # ---- GaussianConditional (reference shapes: y (B,M,1,1) against scales (B,M,4,4)) --------------------
    scales = torch.exp(torch.empty(3, 8, 4, 4).uniform_(np.log(0.05), np.log(64), generator=g))
    y = torch.randn(3, 8, 1, 1, generator=g) * 3
    gc.eval()
    yh, yl = gc(y, scales)
    y2 = torch.randn(3, 8, 4, 4, generator=g) * scales
    y2[0, 0, 0, 0] = 900.0
    yh2, yl2 = gc(y2, scales)
    idx = gc.build_indexes(scales)
    ys = gc.compress(y2, idx)
    fx.update(gc_scales=scales.numpy(), gc_y_bcast=y.numpy(), gc_yhat_bcast=yh.numpy(), gc_lik_bcast=yl.numpy(),
              gc_y=y2.numpy(), gc_yhat=yh2.numpy(), gc_lik=yl2.numpy(), gc_idx=idx.numpy(),
              gc_strings=np.array([np.frombuffer(s, dtype=np.uint8) for s in ys], dtype=object))

    # ---- GDN / IGDN --------------------------------------------------------------------------------------
    for inv in (False, True):
        gdn = R.GDN(10, inverse=inv)
        with torch.no_grad():
            gdn.gamma.add_(torch.rand(10, 10, generator=g) * 0.05)
            gdn.beta.add_(torch.rand(10, generator=g) * 0.5)
        x = torch.randn(2, 10, 6, 5, generator=g)
        tag = "igdn" if inv else "gdn"
        fx.update({f"{tag}_beta": gdn.beta.detach().numpy(), f"{tag}_gamma": gdn.gamma.detach().numpy(),
                   f"{tag}_x": x.numpy(), f"{tag}_y": gdn(x).detach().numpy()})

    # ---- wrappers: RD loss of each model family on a tiny batch ------------------------------------------
    losses = {}
    for kind, tasks, l, c in ((1, ("mono",), 8, 8), (2, ("rgb", "depth_euclidean"), 12, 8),
                              (3, ("rgb", "depth_euclidean", "normal"), 14, 12),
                              (4, ("rgb", "semantic"), 13, 8)):
        torch.manual_seed(100 + kind)
        m = orm.ReferenceCompressor(kind, tasks, l, c, lmbda=1e-2).eval()
        batch = orm.synthetic_batch(tasks, 1, size=256, seed=21)
        with torch.no_grad():
            loss, logs = m.rd_loss(batch, "val")
        losses[str(kind)] = {"loss": float(loss), **{k: float(v) for k, v in logs.items()}}
    json.dump(losses, open(os.path.join(OUT, "wrapper_losses.json"), "w"), indent=1)

    np.savez_compressed(os.path.join(OUT, "rate_path_fixtures.npz"), **fx)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
