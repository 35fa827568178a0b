"""CPU-only: the C-ABI library loads and exports every symbol include/tnb200.h declares."""
import ctypes
import os
import re
from tensornetwork_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
  src = open(os.path.join(ROOT, "include", "tnb200.h")).read()
  return sorted(set(re.findall(r"TNB200_API[^;]*?\b(tnb200_\w+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
  names = _declared()
  assert len(names) >= 25
  lib = ctypes.CDLL(_lib.LIB_PATH)
  for n in names:
    assert hasattr(lib, n), "missing export " + n
    assert n in _lib.SIGNATURES, "no ctypes prototype for " + n
  assert sorted(_lib.SIGNATURES) == names


def test_loads_without_gpu_and_reports_abi():
  lib = _lib.load()
  assert lib.tnb200_abi_version() == 1
  assert ctypes.sizeof(_lib.TensorDesc) == 8 + 4 + 4 + 8 * 16 * 2


def test_import_is_lazy():
  import subprocess, sys
  code = ("import sys; import tensornetwork_b200; "
          "assert 'torch' not in sys.modules, 'torch imported eagerly'")
  subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def test_backend_refuses_cpu():
  import pytest, torch
  if torch.cuda.is_available():
    pytest.skip("GPU present")
  import tensornetwork_b200.backend as b
  with pytest.raises(RuntimeError):
    b.CudaB200Backend()
