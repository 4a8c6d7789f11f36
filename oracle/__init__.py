"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference algorithm for the hot path (MMCTransformer forward, per-video
decode, Gaussian Soft-NMS; plus the value of MMCTransformer.losses).  Nothing under `repurpose_b200/` may import this package; only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` use it, as the
checker / CPU baseline — never as the product path.

Parity pin: the reference ships NO golden vectors, known-answer tests or fixtures for this path
("parity unpinned" by the reference's own tests).  The pin used instead is the
reference ITSELF, imported from /root/reference in the build container by `oracle/make_golden.py`
(and `oracle/make_golden_losses.py`),
which (a) checks this restatement against the reference on the same seeded inputs and (b) writes the
reference's outputs to `tests/golden/*.npz`.  The CPU test-suite re-checks the oracle against those
fixtures, so on the GPU box (where /root/reference does not exist) oracle == reference is already
established when the CUDA path is compared against it.
"""
