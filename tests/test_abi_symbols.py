"""The C-ABI library loads (no GPU needed) and exports every function include/mnf_b200.h declares;
the ctypes binding declares the same set."""

import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "mnf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mnf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from torch_mnf import _lib

    assert os.path.exists(_lib.LIB_PATH), "build first: python torch-mnf_b200/build.py"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    names = header_functions()
    assert len(names) >= 12
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, f"library does not export {missing}"


def test_binding_matches_header():
    import torch_mnf.layers  # noqa: F401  (registers the MNF entry points)
    from torch_mnf import _lib

    assert sorted(_lib.declared_symbols()) == header_functions()
    assert _lib.lib().mnf_abi_version() == _lib.ABI_VERSION == 1


def test_struct_layouts_match_header():
    from torch_mnf import _lib
    from torch_mnf.layers._mnf_ops import KlArgs, KlFusedArgs, RnvpFlow

    assert ctypes.sizeof(_lib.FlowOp) == 4 * (5 + 7 + 2 + 2)  # mnf_flow_op: 16 32-bit words
    assert ctypes.sizeof(RnvpFlow) == 4 * 5 + 4 + 8 * (4 + 4 + 4)  # n_net, sizes[4], pad, 12 pointers
    assert ctypes.sizeof(KlArgs) == 16 + 8 * 14 + 8 + 8 + 8 * 2
    assert ctypes.sizeof(KlFusedArgs) == ctypes.sizeof(KlArgs) + 8 + 8 + 16 + 8 * ctypes.sizeof(RnvpFlow) + 8 * 8 + 8 * 4
    assert ctypes.sizeof(_lib.GatherOut) == 4 + 4 + 8 + 8 * 8 + 8  # mnf_gather_out


def test_argument_errors_do_not_need_a_gpu():
    from torch_mnf import _lib

    lib = _lib.lib()
    ops = (_lib.FlowOp * 1)()
    ops[0].type = 99
    rc = lib.mnf_flow_stack_plan(ops, 1, 2, 16)
    assert rc == -1 and b"bad type" in lib.mnf_last_error()
    rc = lib.mnf_flow_stack_plan(ops, 1, 1000, 16)
    assert rc == -2 and b"dim" in lib.mnf_last_error()
