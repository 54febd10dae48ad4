"""CPU: the C-ABI library loads, exports every symbol include/feastcuda.h declares, its host-only entry points agree with
the oracle, the Python mirror validates arguments like the reference, and the product never routes through oracle/."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import feast_oracle as fo

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib(built_lib):
    import feastcuda as fc
    return fc._lib.load()


def _declared_symbols():
    text = (ROOT / "include" / "feastcuda.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(feastcuda_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    import feastcuda as fc
    names = _declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/feastcuda.h but not exported by libfeastcuda.so"
        assert name in fc._lib.SIGNATURES, f"{name} has no ctypes signature"
    assert set(fc._lib.SIGNATURES) <= set(names)
    assert lib.feastcuda_version() >= 100


def test_struct_layouts_match_the_header(lib):
    """feastcuda_solver_opts / feastcuda_stats are mirrored field by field (sizes per the C layout rules)."""
    import feastcuda as fc
    assert C.sizeof(fc._lib.SolverOpts) == 96
    assert fc._lib.SolverOpts.eps_floor.offset == 80
    hdr = (ROOT / "include" / "feastcuda.h").read_text()
    end = hdr.index("} feastcuda_solver_opts;")
    body = hdr[hdr.rindex("typedef struct {", 0, end):end]
    fields = re.findall(r"^\s*(?:int32_t|double)\s+(\w+)", body, flags=re.M)
    assert fields == [f for f, _ in fc._lib.SolverOpts._fields_]
    sbody = hdr[hdr.index("typedef struct {", hdr.index("} feastcuda_solver_opts;")):hdr.index("} feastcuda_stats;")]
    sfields = []
    for line in sbody.splitlines():
        line = re.sub(r"/\*.*", "", line)
        m = re.match(r"\s*(?:int64_t|double)\s+(.*);", line)
        if m:
            sfields += [re.sub(r"\[.*\]", "", x).strip() for x in m.group(1).split(",")]
    assert sfields == [f for f, _ in fc._lib.Stats._fields_]


def test_feastinit_feastdefault_match_oracle(lib):
    import feastcuda as fc
    assert fc.feastinit() == fo.feastinit()
    cases = [{}, {2: 16}, {2: 24}, {3: 6, 4: 3}, {8: 24}, {16: 1, 2: 30}, {19: 45}, {8: 48}]
    for over in cases:
        a, b = fc.feastinit(), fo.feastinit()
        for k, v in over.items():
            a[k - 1] = v
            b[k - 1] = v
        fc.feastdefault_(a)
        fo.feastdefault(b)
        assert a == b, over
    for bad in ({2: 21}, {3: 17}, {16: 5}, {8: 1}, {8: 41}, {19: 200}):
        a, b = fc.feastinit(), fo.feastinit()
        for k, v in bad.items():
            a[k - 1] = v
            b[k - 1] = v
        with pytest.raises(ValueError):
            fc.feastdefault_(a)
        with pytest.raises(ValueError):
            fo.feastdefault(b)


@pytest.mark.parametrize("ne", [3, 4, 8, 16, 20, 24, 32, 56])
def test_contours_match_oracle(lib, ne):
    import feastcuda as fc
    for (Emin, Emax) in ((0.5, 1.5), (-3.0, 7.25), (0.0, 0.0222)):
        for rule in (0, 1):
            if rule == 0 and ne > 20 and ne not in (24, 32, 40, 48, 56):
                continue
            a, b = fc.feastinit(), fo.feastinit()
            a[1] = b[1] = ne
            a[15] = b[15] = rule
            Z, W = fc.feast_contour(Emin, Emax, a)
            Zo, Wo = fo.feast_contour(Emin, Emax, b)
            assert np.allclose(Z, Zo, rtol=1e-13, atol=1e-15) and np.allclose(W, Wo, rtol=1e-12, atol=1e-16)
    for (Emid, r) in ((0.0, 2.0), (1.0 + 0.5j, 0.3)):
        for rot in (0, 30):
            a, b = fc.feastinit(), fo.feastinit()
            a[7] = b[7] = max(ne, 4)
            a[18] = b[18] = rot
            Z, W = fc.feast_gcontour(Emid, r, a)
            Zo, Wo = fo.feast_gcontour(Emid, r, b)
            assert np.allclose(Z, Zo, rtol=1e-13, atol=1e-15) and np.allclose(W, Wo, rtol=1e-12, atol=1e-16)


def test_node_partition_is_the_reference_block_distribution(lib):
    """parallel/feast_mpi.jl:36-43: contiguous blocks, the first (ne mod P) ranks get one extra node."""
    import feastcuda as fc
    for ne in (1, 8, 16, 24, 7):
        for P in (1, 2, 3, 4, 8):
            owned = []
            for rank in range(P):
                s, c = fc.node_partition(ne, P, rank)
                assert (s, c) == fo.node_partition(ne, P, rank)
                owned += list(range(s, s + c))
            assert owned == list(range(ne))


def test_no_gpu_means_loud_failure_not_a_cpu_fallback(lib):
    import torch
    import feastcuda as fc
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(fc.FeastCudaError) as ei:
        fc.Engine(0)
    assert ei.value.code == fc._lib.ERR_CUDA and "no CPU fallback" in str(ei.value)
    import scipy.sparse as sp
    with pytest.raises(fc.FeastCudaError):
        fc.feast_scsrev(sp.identity(4, format="csc"), 0.0, 2.0, 2, fc.feastinit())


def test_argument_errors_are_raised_host_side_before_any_device_call(lib):
    """interfaces/feast_interfaces.jl:152-155, core/feast_aux.jl:369-399: ArgumentError -> ValueError."""
    import scipy.sparse as sp
    import feastcuda as fc
    A = fo.laplacian_1d(10).tocsc()
    with pytest.raises(ValueError):
        fc.feast_scsrev(A, 1.0, 0.0, 4, fc.feastinit())
    with pytest.raises(ValueError):
        fc.feast_scsrev(A, 0.0, 1.0, 11, fc.feastinit())
    with pytest.raises(ValueError):
        fc.feast(sp.csc_matrix(np.array([[1.0, 2.0], [0.0, 3.0]])), (0.0, 4.0), M0=2)
    with pytest.raises(ValueError):
        fc.feast(np.array([[1.0 + 0j, 2.0 + 1j], [3.0 - 1j, 4.0]]), (0.0, 5.0), M0=2)
    with pytest.raises(ValueError):
        fc.feast_hcsrev(sp.csc_matrix(np.array([[1.0, 2.0j], [2.0j, 3.0]])), 0.0, 4.0, 2, fc.feastinit())
    with pytest.raises(ValueError):
        fc.feast_scsrev(A, 0.0, 1.0, 4, fc.feastinit(), solver="cholesky")
    with pytest.raises(ValueError):
        fc.check_feast_srci_input(10, 4, 0.0, 1.0, [0] * 10)


def test_product_never_imports_the_oracle():
    pkg = ROOT / "feastkit.jl_b200"
    for path in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.hpp")):
        text = path.read_text()
        assert not re.search(r"^\s*(import|from)\s+(feast_oracle|feast_port|oracle)\b", text, flags=re.M), path
        assert "sys.path" not in text or "oracle" not in text, path


def test_rci_init_handshake(lib):
    """runtests.jl:72-118, 1115-1126 (KA14): after INIT info == 0 and ijob == 10 (Feast_RCI_FACTORIZE) for srci/hrci/grci/pdfeast_srci;
    invalid sizes come back as FeastError codes in `info`, not exceptions (kernel/feast_kernel.jl:19-35)."""
    import feastcuda as fc
    N, M0 = 10, 4

    def bufs(cplx_q):
        return dict(work=np.zeros((N, M0)), workc=np.zeros((N, M0), dtype=complex),
                    Aq=np.zeros((M0, M0), dtype=complex if cplx_q else float), Sq=np.zeros((M0, M0), dtype=complex if cplx_q else float),
                    lam=np.zeros(M0, dtype=complex if cplx_q == "g" else float), q=np.zeros((N, M0), dtype=complex if cplx_q else float),
                    res=np.zeros(M0))
    for fn, cplx_q, args in ((fc.feast_srci, False, (0.0, 1.0)), (fc.pdfeast_srci, False, (0.0, 1.0)), (fc.feast_hrci, True, (0.0, 1.0)),
                             (fc.feast_grci, "g", (0.0 + 0.0j, 1.0))):
        b = bufs(cplx_q)
        ijob, Ze, eps, loop, mode, info = fc.Ref(-1), fc.Ref(0j), fc.Ref(0.0), fc.Ref(0), fc.Ref(0), fc.Ref(-1)
        fpm = fc.feastinit()
        st = fn(ijob, N, Ze, b["work"], b["workc"], b["Aq"], b["Sq"], fpm, eps, loop, args[0], args[1], M0, b["lam"], b["q"], mode, b["res"], info)
        assert info.v == 0 and ijob.v == 10 and loop.v == 0
        assert st.initialized and st.ne == (16 if cplx_q == "g" else 8) and abs(Ze.v - st.Zne[0]) == 0
        assert fpm[49] == 1 and fpm[50] == st.ne                      # fpm[50], fpm[51] of the reference
        blk = b["workc"] if cplx_q else b["work"]
        assert np.allclose(np.linalg.norm(blk, axis=0), 1.0)          # unit-norm seeded subspace
    b = bufs(False)
    ijob, info = fc.Ref(-1), fc.Ref(-1)
    fc.feast_srci(ijob, N, fc.Ref(0j), b["work"], b["workc"], b["Aq"], b["Sq"], fc.feastinit(), fc.Ref(0.0), fc.Ref(0), 1.0, 0.0, M0,
                  b["lam"], b["q"], fc.Ref(0), b["res"], info)
    assert info.v == 3 and ijob.v == -1                                # Feast_ERROR_EMIN_EMAX
    info = fc.Ref(-1)
    fc.feast_srci(fc.Ref(-1), N, fc.Ref(0j), b["work"], b["workc"], b["Aq"], b["Sq"], fc.feastinit(), fc.Ref(0.0), fc.Ref(0), 0.0, 1.0, 11,
                  b["lam"], b["q"], fc.Ref(0), b["res"], info)
    assert info.v == 2                                                  # Feast_ERROR_M0
    with pytest.raises(ValueError):
        fc.feast_srci(fc.Ref(99), N, fc.Ref(0j), b["work"], b["workc"], b["Aq"], b["Sq"], fc.feastinit(), fc.Ref(0.0), fc.Ref(0), 0.0, 1.0,
                      M0, b["lam"], b["q"], fc.Ref(0), b["res"], fc.Ref(0), state=fc.FeastRCIState())


def test_complex_symmetric_and_polynomial_host_helpers():
    """Host-side parts of the complex-symmetric / polynomial families (no GPU): symmetry checks raise before any device call
    (runtests.jl:270-273), the symmetric band expands to the general band without conjugation, the companion pencil
    (dense/feast_dense.jl:727-760) has the polynomial's eigenpairs."""
    import scipy.linalg as sla
    import feastcuda as fc
    from feastcuda.families import _symmetric_band_to_general
    with pytest.raises(ValueError, match="complex-symmetric"):
        fc.feast_geev_complex_sym(np.array([[1, 2], [0, 3]], dtype=complex), 1.0 + 0.1j, 1.5, 2, fc.feastinit())
    with pytest.raises(ValueError, match="d\\+1 coefficient"):
        fc.feast_pep([np.eye(2), np.eye(2)], 2, 0j, 1.0, 2, fc.feastinit())
    rng = np.random.default_rng(0)
    n, k = 9, 2
    A = np.zeros((n, n), dtype=complex)
    for d in range(k + 1):
        v = rng.standard_normal(n - d) + 1j * rng.standard_normal(n - d)
        A += np.diag(v, d) + (np.diag(v, -d) if d else 0)
    assert np.array_equal(fo.general_banded_to_full(_symmetric_band_to_general(fo.full_to_banded(A, k), k), k), A)
    N = 4
    K, Cm, M = rng.standard_normal((N, N)), rng.standard_normal((N, N)), np.eye(N) + 0.1 * rng.standard_normal((N, N))
    Al, Bl = fc.companion_linearization([K, Cm, M])
    w, V = sla.eig(Al, Bl)
    for i, lam in enumerate(w):
        assert np.linalg.norm((K + lam * Cm + lam * lam * M) @ V[:N, i]) < 1e-10 * np.linalg.norm(V[:N, i]) * max(1, abs(lam)) ** 2
        assert np.allclose(V[N:, i], lam * V[:N, i], atol=1e-10 * max(1, abs(lam)))


def test_zolotarev_contour_matches_the_oracle(lib):
    """fpm[16] = 2: libfeastcuda's compiled tables against the oracle's (both generated from the reference's ZOLOTAREV_TABLES)."""
    import feastcuda as fc
    for ne in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 16, 20):
        fpm, fpo = fc.feastinit(), fo.feastinit()
        fpm[1] = fpo[1] = ne
        fpm[15] = fpo[15] = 2
        Z, W = fc.feast_contour(0.5, 3.5, fpm)
        Zo, Wo = fo.feast_contour(0.5, 3.5, fpo)
        assert np.array_equal(Z, Zo) and np.array_equal(W, Wo)
