"""The C-ABI library loads without a GPU and exports exactly what include/hge_b200.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from hypergraphembedding_b200 import _native

HEADER = os.path.join(ROOT, "include", "hge_b200.h")


def declared_functions():
  text = open(HEADER).read()
  text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
  return sorted(set(re.findall(r"\b(hge_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_something():
  names = declared_functions()
  assert "hge_algdist_run" in names and "hge_incidence_create" in names


def test_library_exports_every_declared_symbol():
  lib = ctypes.CDLL(_native.LIB_PATH)
  missing = [n for n in declared_functions() if not hasattr(lib, n)]
  assert not missing, "declared in hge_b200.h but not exported: %s" % missing


def test_ctypes_signatures_cover_header():
  declared = set(declared_functions())
  bound = set(_native.SIGNATURES)
  assert bound <= declared, "bound but not declared: %s" % sorted(bound - declared)
  assert declared <= bound, "declared but not bound: %s" % sorted(declared - bound)


def test_version_and_error_string():
  lib = _native.load_library()
  assert lib.hge_version() >= 100
  assert isinstance(_native.last_error(), str)


def test_no_silent_cpu_fallback():
  """Without a CUDA device creating a context must fail loudly (never compute on the CPU)."""
  import torch
  if torch.cuda.is_available():
    pytest.skip("a GPU is present")
  with pytest.raises(_native.NativeError):
    _native.Context(0)


def test_missing_library_message(monkeypatch):
  monkeypatch.setattr(_native, "_lib", None)
  monkeypatch.setattr(_native, "LIB_PATH", "/nonexistent/libhge_b200.so")
  with pytest.raises(_native.NativeLibraryMissing):
    _native.load_library()
