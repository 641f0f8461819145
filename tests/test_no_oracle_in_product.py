"""The oracle is test infrastructure: nothing under torch-mnf_b200/ may import or execute it, and the
product must fail loudly (no CPU fallback) when the CUDA library is missing or tensors are on the CPU."""

import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "torch-mnf_b200")


def test_product_never_references_the_oracle():
    bad = []
    for base, _, files in os.walk(PKG):
        if os.sep + "build" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b|oracle/|/root/reference", text, flags=re.M):
                    bad.append(os.path.join(base, f))
    assert not bad, f"product files referencing the oracle / reference checkout: {bad}"


def test_only_tests_smoke_and_bench_cpu_legs_use_the_oracle():
    """Outside tests/, the oracle may appear only in __graft_entry__.smoke() and in bench.py's CPU legs."""
    bad = []
    for base, dirs, files in os.walk(ROOT):
        dirs[:] = [d for d in dirs if d not in (".git", "tests", "oracle", "gpurun_out", "__pycache__", "build", "baseline")]
        for f in files:
            path = os.path.join(base, f)
            if not f.endswith(".py") or os.path.relpath(path, ROOT) in ("bench.py", "__graft_entry__.py"):
                continue
            if re.search(r"^\s*(from|import)\s+oracle\b", open(path).read(), flags=re.M):
                bad.append(os.path.relpath(path, ROOT))
    assert not bad, f"files outside tests/ importing the oracle: {bad}"
    bench = open(os.path.join(ROOT, "bench.py")).read()
    # every oracle import in bench.py sits inside a function of the CPU legs (no module-level import)
    assert not re.search(r"^(from|import)\s+oracle\b", bench, flags=re.M)


def test_missing_library_fails_loudly():
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import os; os.environ['MNF_B200_LIB'] = '/nonexistent/libmnf_b200.so'\n"
        "import torch, torch_mnf.flows as nf\n"
        "try:\n"
        "    nf.AffineConstantFlow(2).forward(torch.zeros(4, 2, device='cuda' if torch.cuda.is_available() else 'cpu'))\n"
        "except (ImportError, RuntimeError) as e:\n"
        "    print('RAISED', type(e).__name__)\n" % PKG
    )
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "RAISED" in out.stdout, out.stdout + out.stderr


def test_cpu_tensors_are_rejected():
    import torch_mnf.flows as nf
    from torch_mnf.layers import MNFLinear

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        nf.AffineHalfFlow(2, parity=False).forward(torch.zeros(4, 2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        MNFLinear(4, 3).forward(torch.zeros(2, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        MNFLinear(4, 3).kl_div()
