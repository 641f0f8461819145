"""CPU oracle for the torch-mnf hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU (PyTorch fp32, ATen) restatement of the reference's
algorithm for the flow forward/inverse + log-det path and the MNF layer
forward / kl_div path.  It is the *checker* the CUDA kernels are compared
with; it is never the thing shipped or measured as the product:

* only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
  ``cpu_baseline`` / ``--impl reference`` legs may import it;
* nothing under ``torch-mnf_b200/`` imports it (``tests/test_no_oracle_in_product.py``
  greps for that), and the product raises when the CUDA library is missing.

The reference is pure Python on top of ATen, so the restatement uses the same
ATen CPU ops in the same order as the reference call sites (cited per function
as ``file:line`` relative to the reference checkout).  That keeps the oracle
within a few ulp of the reference and makes it a fair CPU baseline ("port").

Parity pin: ``oracle/make_golden.py`` imports the real reference from
``/root/reference`` (build container only), records inputs, weights, injected
noise and outputs, and writes them to ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` checks this oracle against those vectors
(bit-for-bit where the op order is identical, 1e-6 otherwise), so the oracle
is pinned to outputs of the reference itself (torch 2.11 CPU).
"""

from . import flows_cpu, mnf_cpu, noise  # noqa: F401
