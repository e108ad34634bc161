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


# ---- the ctypes mirror follows the header: argument counts and struct layouts --------------------
def _prototypes():
    """name -> number of parameters, from the declarations of include/abnet3_b200.h."""
    text = open(os.path.join(ROOT, "include", "abnet3_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"ABN_API\s+[\w\s\*]+?\b(abn_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_ctypes_argument_counts_match_the_prototypes():
    protos = _prototypes()
    assert set(protos) == set(_lib.SIGNATURES)
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        assert len(argtypes) == protos[name], "%s: %d ctypes arguments, %d in the header" % (
            name, len(argtypes), protos[name])


STRUCTS = {"abn_dropout": "Dropout", "abn_gemm_problem": "GemmProblem", "abn_mlp_layer": "MlpLayer",
           "abn_mlp_loss": "MlpLoss", "abn_mlp_dlayer": "MlpDLayer", "abn_param_segment": "ParamSegment",
           "abn_dp_peers": "DpPeers", "abn_dp_push": "DpPush"}


def test_ctypes_structures_have_the_layout_of_the_c_structs(tmp_path):
    """The header compiles as plain C (gcc); every struct the ABI passes by pointer has the same size
    and field offsets as its ctypes mirror in abnet3_b200/_lib.py."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    text = open(os.path.join(ROOT, "include", "abnet3_b200.h")).read()
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "abnet3_b200.h"', "int main(void) {"]
    fields = {}
    for c_name, py_name in STRUCTS.items():
        cls = getattr(_lib, py_name)
        fields[c_name] = [f[0] for f in cls._fields_]
        end = re.search(r"\}\s*%s\s*;" % c_name, text).start()
        beg = text.rfind("typedef struct", 0, end)
        body = text[text.index("{", beg) + 1:end]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        c_fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                c_fields.append(re.sub(r"\[.*?\]", "", part).replace("*", " ").split()[-1])
        assert len(c_fields) == len(fields[c_name]), (c_name, c_fields, fields[c_name])
        lines.append('printf("%s %%zu", sizeof(%s));' % (c_name, c_name))
        for f in c_fields:
            lines.append('printf(" %%zu", offsetof(%s, %s));' % (c_name, f))
        lines.append('printf("\\n");')
    lines += ["return 0;", "}"]
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    for line in out.strip().splitlines():
        parts = line.split()
        c_name, size, offs = parts[0], int(parts[1]), [int(x) for x in parts[2:]]
        cls = getattr(_lib, STRUCTS[c_name])
        assert ctypes.sizeof(cls) == size, c_name
        assert [getattr(cls, f).offset for f in fields[c_name]] == offs, c_name
