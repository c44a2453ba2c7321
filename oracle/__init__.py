"""CPU oracle for the causal-OT loss path.  TEST INFRASTRUCTURE — never imported by the product
package `kccotgan_b200`; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU legs use it.

Pinning status: the reference (neuripss2020/kccotgan) ships no tests, fixtures or known-answer
vectors for this path, and TensorFlow — the library its arithmetic runs on —
is absent from this image.  The oracle is therefore pinned against OUTPUTS OF THE REFERENCE'S OWN
SOURCE, executed in this container through `oracle/tf_shim` (a torch-backed module named
`tensorflow`) by `oracle/ref_exec.py`; the frozen results live in `tests/golden/*.npz`, generated
by `oracle/make_golden.py`.  On the GPU box `/root/reference` does not exist; there only
`oracle/closed_form.py` (numpy fp64) and `oracle/port_torch.py` (reference formulation on torch
CPU) are used, both of which are checked against the golden vectors by `tests/test_oracle.py`.
"""
