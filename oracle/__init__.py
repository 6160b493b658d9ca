"""ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.

CPU restatement of the rate path of narekvslife/multi-modal-neural-compression:

  rans_cdf_ref.c        plain C: pmf_to_quantized_cdf + rANS encode/decode (CompressAI 1.2.4 C++ modules)
  native.py             ctypes view of the C oracle, with CompressAI's pybind11 signatures
  compressai_ref.py     stock-torch restatement of EntropyBottleneck / GaussianConditional / GDN / LowerBound /
                        ScaleHyperprior in CompressAI's own op order
  reference_models.py   restatement of the reference's four multi-task wrappers and their RD-loss formulas
                        (/root/reference/src/models/*.py, src/loss_balancing.py)

Who may import this package: `tests/`, `__graft_entry__.smoke()`, and the `cpu_baseline` / `--impl reference`
legs of `bench.py` — there only as the checker or the CPU baseline, never as the thing shipped.  Nothing under
`multi-modal-neural-compression_b200/` imports it; the product fails loudly when its CUDA library is missing.

Why "unpinned": the arithmetic lives in `compressai==1.2.4` (/root/reference/requirements.txt:10), which is not
vendored in /root/reference, not installed here, and not downloadable; the reference has no tests or golden
vectors for this path.  The only pins are self-derived known answers checked in
tests/test_oracle_known_answers.py and the fixtures under tests/golden/ generated from this oracle.
"""
