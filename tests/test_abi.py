"""CPU checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol that
include/rbc3d.h declares, its host-side arithmetic (SetEwaldPrms, EwaldCoeff_*_Exact) matches the oracle, and the
product path fails loudly -- no CPU fallback -- when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from rbc3d_b200 import capi, ewald

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rbc3d.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rbc3d_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from rbc3d_b200 import build
    build.build()
    lib = C.CDLL(capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rbc3d.h but not exported"
        assert n in capi.SIGNATURES, f"{n} has no ctypes signature in rbc3d_b200/capi.py"
    assert set(capi.SIGNATURES) <= set(names)


@pytest.mark.parametrize("Lb,nranks", [([10.5, 10.5, 8.0], 1), ([10.5, 10.5, 8.0], 2), ([10.5, 10.5, 11.43], 1),
                                       ([10.5, 10.5, 30.0], 1), ([57.2588] * 3, 1), ([57.2588] * 3, 8), ([3.0, 3.0, 3.0], 1)])
def test_set_ewald_prms_matches_oracle(oracle_lib, Lb, nranks):
    rc, Nb = ewald.SetEwaldPrms(Lb, 0.44, 1e-3, 8, nranks)
    orc = oracle_lib.Oracle(Lb, nranks=nranks)
    assert rc == orc.rc and Nb == orc.Nb
    assert Nb[2] % nranks == 0 and Nb[2] >= nranks * 8          # ModConf.F90:394-395


def test_baseline_table_values():
    """the derived parameters BASELINE.md / SURVEY.md section 8 quote for the named configs"""
    rc, Nb = ewald.SetEwaldPrms([10.5, 10.5, 8.0])
    assert abs(rc - 1.1986) < 1e-4 and Nb == [48, 48, 36]
    assert ewald.SetEwaldPrms([10.5, 10.5, 11.43])[1] == [48, 48, 52]
    assert ewald.SetEwaldPrms([10.5, 10.5, 30.0])[1] == [48, 48, 136]
    assert ewald.SetEwaldPrms([57.2588] * 3)[1] == [256, 256, 256]
    rc3, _ = ewald.SetEwaldPrms([3.0, 3.0, 3.0])
    assert abs(rc3 - 3.0 / 3.001) < 1e-15                       # rc = min(rc, min(Lb)/3.001)


def test_exact_coefficients_match_oracle(oracle_lib):
    for r in np.linspace(0.01, 2.0, 60):
        for alpha in (0.44, 0.1, -1.0):
            assert ewald.EwaldCoeff_SL_Exact(r, alpha) == oracle_lib.Oracle.ewald_sl_exact(r, alpha)
            assert ewald.EwaldCoeff_DL_Exact(r, alpha) == oracle_lib.Oracle.ewald_dl_exact(r, alpha)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(capi.Rbc3dError):
        ewald.EwaldOperator([10.5, 10.5, 8.0])


def test_product_does_not_import_oracle():
    """the oracle is test infrastructure: nothing under rbc3d_b200/ may reference it"""
    pkg = os.path.join(ROOT, "rbc3d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "rbc3d_oracle" not in txt, f
