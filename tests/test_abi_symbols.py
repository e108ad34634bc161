"""CPU: the C-ABI library loads and exports every symbol include/*.h declares.
No compute call is made (there is no GPU here and no CPU path in the library)."""
import ctypes
import os
import re

import pytest

from abnet3_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def so_path():
    return build.build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "abnet3_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(abn_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for name in ("abn_align_pairs", "abn_cosine_distance", "abn_dtw_from_dist",
                 "abn_pair_loss", "abn_linear_forward", "abn_linear_backward",
                 "abn_optimizer_step", "abn_gather_batch", "abn_diff_pairs"):
        assert name in syms


def test_library_exports_every_declared_symbol(so_path):
    handle = ctypes.CDLL(so_path)
    for name in declared_symbols():
        assert hasattr(handle, name), name


def test_ctypes_signatures_cover_the_header(so_path):
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.lib()
    assert lib.abn_version() >= 100


def test_entry_points_refuse_to_run_without_sm100(so_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.lib()
    rc = lib.abn_align_pairs(None, 0, 280, None, 1, 80, 0, None, None, None, None, None, None,
                             None, 0, None)
    assert rc == 38          # ABN_ENOSYS: no CPU fallback by design
    assert b"no" in lib.abn_last_error().lower()
