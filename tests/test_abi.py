"""CPU: the C-ABI library builds, loads without a GPU and exports every symbol that
include/gkn_b200.h declares (no compute calls here)."""
import ctypes
import re
from pathlib import Path

import pytest

from molecular_dynamics_neural_operator_b200 import _build, _lib

HEADER = Path(__file__).resolve().parents[1] / "include" / "gkn_b200.h"


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(gkn_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for required in ("gkn_contract_fwd_bf16", "gkn_contract_dx_bf16", "gkn_scatter_mean_fwd",
                     "gkn_radius_graph_count", "gkn_radius_graph_fill", "gkn_graph_prepare",
                     "gkn_last_error", "gkn_version"):
        assert required in names


def test_library_builds_and_exports_every_declared_symbol():
    path = _build.build()
    assert path.exists()
    lib = ctypes.CDLL(str(path))
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in gkn_b200.h but not exported"
    # and the Python binding table covers exactly the header
    assert sorted(_lib.SIGNATURES) == declared_functions()


def test_version_and_error_string_without_gpu():
    lib = _lib.load()
    assert lib.gkn_version() == 100
    assert isinstance(_lib.last_error(), str)
    assert lib.gkn_graph_prepare_workspace_bytes(10, 100) >= (2 * 10 + 2 * 100) * 4
    assert lib.gkn_radius_graph_workspace_bytes(2, 504) > 0


def test_cpu_tensors_are_rejected_loudly():
    import torch

    from molecular_dynamics_neural_operator_b200 import ops

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.prepare_graph(torch.zeros(2, 3, dtype=torch.long), 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.radius_graph(torch.zeros(1, 4, 3), 8.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.contract_fwd(torch.zeros(4, 64), torch.zeros(1024, 64), torch.zeros(4, dtype=torch.int32),
                         torch.zeros(3, 32), "fp32")


def test_entry_points_reject_null_pointers_without_touching_the_gpu():
    """Argument validation happens before any CUDA call, so it is testable on a CPU-only box."""
    lib = _lib.load()
    rc = lib.gkn_contract_fwd_bf16(None, None, None, None, None, 128, 64, 64, None)
    assert rc != 0 and "NULL" in _lib.last_error()
    rc = lib.gkn_contract_fwd_bf16(None, None, None, None, None, 0, 64, 64, None)
    assert rc == 0  # empty edge set is a no-op
    rc = lib.gkn_radius_graph_count(None, 1, 10, 8.0, None, None, None, 1 << 20, None)
    assert rc != 0 and "NULL" in _lib.last_error()
    rc = lib.gkn_segment_matvec(None, None, None, None, None, None, None, None, None, None, None, None,
                                5, 64, 0, None)
    assert rc != 0
