"""CPU oracle for the AVID / AVID-CMA training hot path.

TEST INFRASTRUCTURE ONLY.  This package is a plain torch-CPU / numpy restatement of
the reference algorithm (facebookresearch/AVID-CMA, criterions/{nce,avid,avid_cma}.py
and models/{video,audio,network_blocks,av_wrapper}.py).  It may be imported only by
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py`.  Nothing under `avid_cma_b200/` imports it: the product path has no CPU
fallback and fails loudly when the CUDA library is missing.

Parity pin: the reference ships no tests or golden vectors.  The oracle
is pinned against outputs of the reference itself, imported read-only from
/root/reference in the build container by `tests/golden/make_golden.py`; the resulting
vectors are committed under `tests/golden/*.npz` and checked by
`tests/test_oracle_golden.py` (no GPU needed).
"""
