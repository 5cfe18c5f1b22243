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


def test_model_dump_load_roundtrip(tmp_path):
    """dump / load text format (factorization_machine.nim:142-220,
    field_aware_factorization_machine.nim:95-158): every parameter survives bit-exactly, the header
    lines are the reference's (a file written by nimfm parses, and vice versa)."""
    import nimfm_b200 as nf
    rng = np.random.default_rng(3)
    for fit_lower, fit_linear in ((nf.explicit, True), (nf.augment, False), (nf.none, True)):
        fm = nf.newFactorizationMachine(nf.classification, degree=3, nComponents=4, fitLower=fit_lower,
                                        fitLinear=fit_linear, randomState=7, scale=0.05)
        fm.P = rng.standard_normal((fm.nOrders, 4, 6 + fm.nAugments)) * 1e-3
        fm.w, fm.intercept, fm.isInitialized = rng.standard_normal(6), -1.5e-7, True
        p = str(tmp_path / "fm.txt")
        fm.dump(p)
        lines = open(p).read().splitlines()
        assert lines[:9] == ["task: c", "nFeatures: 6", "degree: 3", "nComponents: 4", f"fitLower: {fit_lower}",
                             "fitIntercept: true", f"fitLinear: {'true' if fit_linear else 'false'}",
                             "randomState: 7", "scale: 0.05"]
        assert lines[9] == "lams:" and lines[11] == "P[0]:" and lines[-3] == "w:" and lines[-1].startswith("intercept: ")
        g = nf.FactorizationMachine.load(p, warmStart=True)
        assert g.isInitialized and g.warmStart and g.degree == 3 and g.fitLower == fit_lower
        assert np.array_equal(g.P, fm.P) and np.array_equal(g.w, fm.w) and g.intercept == fm.intercept
        assert np.array_equal(g.lams, fm.lams)
    with pytest.raises(nf.NotFittedError):
        nf.newFactorizationMachine(nf.regression).dump(str(tmp_path / "x"))
    ffm = nf.newFieldAwareFactorizationMachine(nf.regression, nComponents=3, fitLinear=False)
    ffm.P, ffm.w, ffm.intercept, ffm.isInitialized = rng.standard_normal((2, 5, 3)), rng.standard_normal(5), 0.5, True
    p = str(tmp_path / "ffm.txt")
    ffm.dump(p)
    assert open(p).read().splitlines()[:4] == ["task: r", "nFields: 2", "nFeatures: 5", "nComponents: 3"]
    h = nf.FieldAwareFactorizationMachine.load(p)
    assert np.array_equal(h.P, ffm.P) and np.array_equal(h.w, ffm.w) and h.intercept == 0.5 and not h.fitLinear


def test_oracle_vstack_matches_dense_definition():
    from oracle import oracle as orc
    from oracle.oracle import CSR
    rng = np.random.default_rng(5)
    mats = [rng.random((n, 7)) * (rng.random((n, 7)) < 0.4) for n in (4, 9, 1)]
    parts = [CSR.from_dense(M) for M in mats]
    st = orc.csr_vstack(parts)
    assert np.array_equal(st.to_dense(), np.vstack(mats)) and st.n == 14
    orc.build()
    cst = orc.csc_vstack([orc.csr_to_csc(p) for p in parts])
    whole = orc.csr_to_csc(st)
    assert np.array_equal(cst.indptr, whole.indptr) and np.array_equal(cst.indices, whole.indices)
    assert np.array_equal(cst.data, whole.data)


def test_nim_binding_declares_every_header_symbol():
    """nim/nimfm_cuda.nim (the {.importc, dynlib.} stub a nimfm maintainer adds) binds
    exactly the symbols include/nimfm_cuda.h declares"""
    nim = open(os.path.join(ROOT, "nim", "nimfm_cuda.nim")).read()
    bound = set(re.findall(r"proc (nimfm_\w+)", nim))
    assert sorted(bound) == header_symbols()
    assert '{.push importc, dynlib: libName, cdecl.}' in nim
