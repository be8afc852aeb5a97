"""CPU-only checks of the consumers' host logic: Saltelli design layout, history-matching bookkeeping, index helpers,
diagnostics helpers, metric / inference plumbing (the GPU parts are covered in test_gpu_consumers.py)."""
import numpy as np
import pytest
from scipy import linalg

from gperks_b200.gp.data.dataset import Dataset
from gperks_b200.perks import diagnostics as dg
from gperks_b200.perks.gsa import SobolGSA
from gperks_b200.perks.history_matching import Wave
from gperks_b200.utils.indices import diff, filter_zscore, part_and_select, whereq_whernot
from gperks_b200.utils.jsonfiles import load_json, save_json
from gperks_b200.utils.test_functions_gsa import Ishigami_theoretical_Si, SobolGstar_theoretical_Si


def test_saltelli_design_layout_matches_reference_stacking():
    """rows = [A | B | for j: for p: A_j with coordinate p taken from B_j]."""
    rng = np.random.default_rng(0)
    ds = Dataset(rng.random((10, 3)), rng.random(10), l_bounds=[0.0, -1.0, 2.0], u_bounds=[1.0, 1.0, 5.0])
    gsa = SobolGSA(ds, n=8, seed=3)
    A, B, X = gsa.saltelli_design()
    n, d = 8, 3
    assert A.shape == (d, n) and B.shape == (d, n) and X.shape == (n * (d + 2), d)
    np.testing.assert_array_equal(X[:n], A.T)
    np.testing.assert_array_equal(X[n:2 * n], B.T)
    for j in range(n):
        for p in range(d):
            row = X[2 * n + j * d + p]
            expect = A[:, j].copy()
            expect[p] = B[p, j]
            np.testing.assert_array_equal(row, expect)
    assert (X >= np.array([0.0, -1.0, 2.0])).all() and (X <= np.array([1.0, 1.0, 5.0])).all()
    A2, B2, X2 = SobolGSA(ds, n=8, seed=3).saltelli_design()
    np.testing.assert_array_equal(X, X2)                       # seeded => reproducible


def test_gsa_thresholding_and_dataframe():
    ds = Dataset(np.random.default_rng(1).random((6, 2)), np.zeros(6))
    gsa = SobolGSA(ds, n=4, seed=0)
    gsa.ST = np.array([[0.5, 0.004], [0.6, 0.02], [0.55, 0.003], [0.58, 0.001]])
    gsa.S1 = np.array([[0.4, 0.002], [0.5, 0.001], [0.45, 0.003], [0.48, 0.02]])
    gsa.correct_Sobol_indices(threshold=0.01)
    assert (gsa.ST[:, 1] == 0).all() and (gsa.S1[:, 1] == 0).all() and (gsa.ST[:, 0] > 0).all()
    gsa.assemble_dataframe()
    assert len(gsa.df) == 2 * 2 * 4 and set(gsa.df["Index"]) == {"ST", "S1"}


def test_theoretical_indices_tables():
    ST, S1, S12 = Ishigami_theoretical_Si()
    assert S1.loc["X1", "Si"] == pytest.approx(0.313905, abs=1e-5) and S1.loc["X3", "Si"] == 0
    assert ST.loc["X3", "STi"] == pytest.approx(0.243684, abs=1e-5)
    ST, S1, _ = SobolGstar_theoretical_Si([0, 1, 4.5], [0.1, 0.2, 0.3], [1.0, 1.0, 1.0])
    assert S1["Si"].sum() <= 1.0 and (ST["STi"].values >= S1["Si"].values).all()


class _StubWave(Wave):
    def compute_impl(self, X):
        return np.abs(X[:, 0] - 0.5) * 10.0, X[:, 1]


def test_wave_bookkeeping_and_json_round_trip(tmp_path):
    X = np.random.default_rng(2).random((200, 2))
    W = _StubWave(emulator=[None], Itrain=np.array([[0.0, 1.0], [0.0, 1.0]]), cutoff=3.0, maxno=1, mean=[0.0], var=[1.0])
    W.find_regions(X)
    assert sorted(W.nimp_idx + W.imp_idx) == list(range(200))
    assert (W.I[W.nimp_idx] < 3.0).all() and (W.I[W.imp_idx] >= 3.0).all()
    np.testing.assert_array_equal(W.reconstruct_tests(), X)
    sel, rest = W.get_nimps(10)
    assert sel.shape == (10, 2) and sel.shape[0] + rest.shape[0] == W.NIMP.shape[0]
    np.random.seed(0)
    n0 = len(W.nimp_idx)
    W2 = W.copy()
    W2.augment_nimp(n0 + 30)
    assert W2.NIMP.shape[0] == n0 + 30 and (np.abs(W2.NIMP[:, 0] - 0.5) * 10.0 < 3.0).all()
    assert ((W2.NIMP > 0.0) & (W2.NIMP < 1.0)).all()
    assert len(W.nimp_idx) == n0                                 # the copy did not touch the original
    W.save(tmp_path / "wave.json")
    W3 = _StubWave()
    W3.load(tmp_path / "wave.json")
    np.testing.assert_allclose(W3.I, W.I)
    assert float(W3.cutoff) == 3.0
    with pytest.raises(ValueError):
        W.get_trains(X, 1000)
    assert W.get_trains(X, 5).shape == (5, 2)


def test_index_helpers():
    P = np.random.default_rng(3).random((64, 3))
    parts, picks = part_and_select(P, 8)
    assert len(parts) == 8 and picks.shape == (8, 3) and sum(len(c) for c in parts) == 64
    hit, rest = whereq_whernot(P, picks)
    np.testing.assert_array_equal(P[hit], picks)
    assert sorted(hit + rest) == list(range(64))
    assert sorted(diff([1, 2, 3, 4], [2, 4])) == [1, 3]
    X = np.vstack([np.zeros((20, 2)) + np.random.default_rng(4).normal(0, 0.1, (20, 2)), [[50.0, 0.0]]])
    keep, out = filter_zscore(X, 3.0)
    assert out == [20] and len(keep) == 20
    d = {"a": np.arange(3), "b": 2.5, "c": [1, 2]}
    save_json(d, "/tmp/_gpx_json_test.json")
    back = load_json("/tmp/_gpx_json_test.json")
    np.testing.assert_array_equal(back["a"], [0, 1, 2])
    assert back["b"] == 2.5


def test_diagnostics_helpers():
    rng = np.random.default_rng(5)
    m = 12
    B = rng.random((m, m))
    C = B @ B.T + m * np.eye(m)
    y_true, y_mean = rng.random(m), rng.random(m)
    std = np.sqrt(np.diag(C))
    e = y_true - y_mean
    assert dg.DChi2(y_true, y_mean, std) == pytest.approx(np.sum((e / std) ** 2))
    assert float(dg.DMD(y_true, y_mean, C)) == pytest.approx(float(e @ np.linalg.solve(C, e)))
    G = dg.DG(y_true, y_mean, C)
    assert float(G @ G) == pytest.approx(float(e @ np.linalg.solve(C, e)))     # whitened errors: ||G||^2 = MD
    L, P = dg.pivoted_Cholesky_factorization(C)
    np.testing.assert_allclose(P.T @ C @ P, L @ L.T, rtol=1e-10, atol=1e-10)
    assert dg.DPC(y_true, y_mean, C).shape == (m,)
    assert dg.DCI(y_true, y_true, std) == 1.0
