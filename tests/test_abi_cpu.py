"""CPU-side checks of the drop-in boundary: the in-tree libnimfm_cuda.so loads, exports every symbol
include/nimfm_cuda.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    hdr = open(os.path.join(ROOT, "include", "nimfm_cuda.h")).read()
    names = re.findall(r"\n(?:int32_t|int64_t|const char \*)\s*\*?(nimfm_\w+)\(", hdr)
    return sorted(set(names))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from nimfm_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(built):
    lib = built.load()
    declared = header_symbols()
    assert len(declared) >= 45
    for name in declared:
        assert hasattr(lib, name), f"libnimfm_cuda.so does not export {name}"
    # the ctypes table and the header agree
    assert sorted(built.SYMBOLS) == declared


def test_header_cites_reference_lines():
    hdr = open(os.path.join(ROOT, "include", "nimfm_cuda.h")).read()
    for cite in ("factorization_machine.nim:100-122", "minibatch_psgd.nim", "adagrad.nim", "cd.nim",
                 "sgd_ffm.nim:11-30", "field_aware_factorization_machine.nim"):
        assert cite in hdr


def test_version_and_no_cpu_fallback(built):
    import torch
    lib = built.load()
    assert lib.nimfm_version() >= 100
    if torch.cuda.is_available():
        pytest.skip("GPU present: the no-device error path is not reachable")
    built.destroy_ctx()
    with pytest.raises(built.NimfmCudaError) as e:
        built.ctx()
    assert "no CPU fallback" in str(e.value)
    import nimfm_b200 as nf
    fm = nf.newFactorizationMachine(nf.regression, degree=2, nComponents=2)
    fm.P, fm.w, fm.isInitialized = np.zeros((1, 2, 3)), np.zeros(3), True
    ds = nf.newCSRDataset([1.0], [0], [0, 1], 1, 3)
    with pytest.raises(built.NimfmCudaError):
        fm.decisionFunction(ds)


def test_host_side_shape_rules():
    """nAugments / nOrders (factorization_machine.nim:81-97), MBPSGD sizes (minibatch_psgd.nim:157-165)"""
    import nimfm_b200 as nf
    for degree in range(2, 6):
        for fl in (True, False):
            fm = nf.newFactorizationMachine(nf.regression, degree=degree, fitLower=nf.augment, fitLinear=fl)
            assert fm.nAugments == (degree - 2 if fl else degree - 1) and fm.nOrders == 1
            fm = nf.newFactorizationMachine(nf.regression, degree=degree, fitLower=nf.explicit, fitLinear=fl)
            assert fm.nAugments == 0 and fm.nOrders == degree - 1
            fm = nf.newFactorizationMachine(nf.regression, degree=degree, fitLower=nf.none)
            assert fm.nAugments == 0 and fm.nOrders == 1
    with pytest.raises(ValueError):
        nf.newFactorizationMachine(nf.regression, degree=0)
    with pytest.raises(ValueError):
        nf.newFactorizationMachine(nf.regression, nComponents=0)
    ds = nf.newCSRDataset(np.ones(6), [0, 1, 2, 0, 1, 2], [0, 3, 6], 2, 1000)
    opt = nf.newMBPSGD()
    assert opt.resolve_sizes(ds) == ((1000 * 2) // 6, 1)
    opt = nf.newMBPSGD(miniBatchSize=1)
    assert opt.resolve_sizes(ds) == (1, 2)
    y = nf.newFactorizationMachine(nf.classification).checkTarget([0.5, -2.0, 0.0])
    assert list(y) == [1.0, -1.0, 0.0]            # fm_base.nim:29-36 sgn
    with pytest.raises(ValueError):
        nf.newCSRDataset([1.0], [0], [0, 1, 1], 1, 3)
