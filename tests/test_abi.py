"""The C-ABI library loads without a GPU and exports every symbol that
include/bambu_em.h declares; without a device the calls fail loudly."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "bambu_em.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bambu_em\w*|bambu_emWithL1|bambu_quant\w*|bambu_wire\w*|bambu_equirc\w*)\s*\(", src)) - {"bambu_em_plan", "bambu_equirc"})


@pytest.fixture(scope="module")
def built_lib():
    from bambu_b200 import _lib
    _lib.build()
    return _lib


def test_every_declared_symbol_is_exported_and_bound(built_lib):
    L = built_lib.lib()
    names = _declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/bambu_em.h but not exported"
        assert n in built_lib.SYMBOLS, f"{n} has no ctypes prototype in bambu_b200/_lib.py"
    assert sorted(built_lib.SYMBOLS) == names
    assert b"sm_100a" in L.bambu_em_version()


def test_no_cpu_fallback(built_lib):
    """Without a B200 the product path must raise, not compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from bambu_b200.em import abundance_quantification
    from bambu_b200 import synthetic
    a = synthetic.config2(scale=0.002)
    with pytest.raises(built_lib.BambuEmError) as e:
        abundance_quantification(a)
    assert e.value.code == built_lib.ENODEVICE


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under bambu_b200/ may reference it."""
    pkg = os.path.join(ROOT, "bambu_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "em_oracle" not in txt and "libem_oracle" not in txt, f


def test_synthetic_config2_shape():
    from bambu_b200 import synthetic
    a = synthetic.config2(scale=0.02)
    a.validate()
    M, N, e = a.group_shapes()
    assert a.n_rows == 5000 and a.n_cols == 10000
    assert abs(a.nnz - 40000) < 4000
    # every transcript and every gene keeps an observation (nothing would be filtered)
    live = a.Y != 0
    rows_of = np.repeat(np.arange(a.n_rows), np.diff(a.row_nnz_ptr))
    grp_of_row = np.repeat(np.arange(a.n_groups), M)
    col_global = a.grp_col_ptr[grp_of_row[rows_of]] + a.col_idx
    assert np.all(np.bincount(rows_of, weights=live[col_global], minlength=a.n_rows) > 0)


def test_r_shim_typechecks(built_lib):
    """r/b200_shim.c (the .Call binding a bambu maintainer would add) cannot be built here -- no R --
    but it can be type-checked against include/bambu_em.h with stand-in R headers, and every library symbol it
    calls must be exported."""
    import shutil
    import subprocess
    shim = os.path.join(ROOT, "r", "b200_shim.c")
    src = re.sub(r"/\*.*?\*/", "", open(shim).read(), flags=re.S)
    used = set(re.findall(r"\b(bambu_(?:em|emWithL1|quant|wire)\w*)\s*\(", src))
    assert {"bambu_em_batched", "bambu_em_batched_v2", "bambu_wire_create", "bambu_emWithL1", "bambu_quantDT_dotC",
            "bambu_em_last_error", "bambu_em_shutdown"} <= used
    L = built_lib.lib()
    for n in used:
        assert n in built_lib.SYMBOLS and hasattr(L, n), n
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-Wno-cast-function-type", "-fsyntax-only",
                        "-I", os.path.join(ROOT, "tests", "r_stub"), "-I", os.path.join(ROOT, "include"), shim],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
