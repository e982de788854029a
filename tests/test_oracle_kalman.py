"""Known-answer pinning of the oracle's Kalman restatements (oracle/kalman.py, oracle/deepocsort.py::XYSRFilter)
against the UNMODIFIED reference classes: the inputs of the reference's own unit tests
(tests/unit/test_kalman_filters_modes.py, which assert only shapes and finiteness) and a seeded batch with NSA
confidences.  Goldens: tests/golden/make_kalman_golden.py."""
from pathlib import Path

import numpy as np
import pytest

from oracle import kalman as okf
from tests.golden.make_kalman_golden import REF_CASES, XYSR_CASE, seeded_batch

G = np.load(Path(__file__).parent / "golden" / "kalman_reference.npz")
TOL = dict(rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("kind", ["xywh", "xyah"])
def test_reference_unit_test_inputs(kind):
    z0, z1 = REF_CASES[kind]
    m, c = okf.initiate(kind, z0)
    assert np.array_equal(m, G[f"{kind}_ref_init_mean"]) and np.array_equal(c, G[f"{kind}_ref_init_cov"])
    m, c = okf.multi_predict(kind, m[None], c[None])
    np.testing.assert_allclose(m[0], G[f"{kind}_ref_pred_mean"], **TOL)
    np.testing.assert_allclose(c[0], G[f"{kind}_ref_pred_cov"], **TOL)
    m2, c2 = okf.update(kind, m[0], c[0], z1, confidence=0.9)
    np.testing.assert_allclose(m2, G[f"{kind}_ref_upd_mean"], **TOL)
    np.testing.assert_allclose(c2, G[f"{kind}_ref_upd_cov"], **TOL)
    # squared Mahalanobis gate of the updated state (base.py:523-551), as StrongSORT's gate uses it
    pm, pc = okf.project(kind, m2, c2)
    d = z1 - pm
    np.testing.assert_allclose(d @ np.linalg.solve(pc, d), G[f"{kind}_ref_gate"][0], rtol=1e-9)


@pytest.mark.parametrize("kind", ["xywh", "xyah"])
def test_seeded_batch_with_nsa_confidences(kind):
    z0, steps = seeded_batch(kind)
    means, covs = map(np.asarray, zip(*[okf.initiate(kind, z) for z in z0]))
    assert np.array_equal(means, G[f"{kind}_b_init_mean"]) and np.array_equal(covs, G[f"{kind}_b_init_cov"])
    z = z0.copy()
    for r, (dz, conf) in enumerate(steps):
        means, covs = okf.multi_predict(kind, means, covs)
        np.testing.assert_allclose(means, G[f"{kind}_b_pred{r}_mean"], **TOL)
        np.testing.assert_allclose(covs, G[f"{kind}_b_pred{r}_cov"], rtol=1e-12, atol=1e-9)
        z = z + dz
        upd = [okf.update(kind, means[i], covs[i], z[i], confidence=float(conf[i]) if r == 2 else 0.0)
               for i in range(len(z))]
        means, covs = np.asarray([u[0] for u in upd]), np.asarray([u[1] for u in upd])
        np.testing.assert_allclose(means, G[f"{kind}_b_upd{r}_mean"], rtol=1e-10, atol=1e-9)
        np.testing.assert_allclose(covs, G[f"{kind}_b_upd{r}_cov"], rtol=1e-9, atol=1e-9)


def test_xysr_reference_unit_test_inputs():
    from oracle.deepocsort import XYSRFilter

    z0, z1 = XYSR_CASE
    f = XYSRFilter(z0)
    f.x, f.P = G["xysr_ref_init_x"].copy(), G["xysr_ref_init_P"].copy()   # the raw filter's own initiate()
    f.Q, f.R = G["xysr_ref_Q"].copy(), G["xysr_ref_R"].copy()
    f.predict()
    np.testing.assert_allclose(f.x, G["xysr_ref_pred_x"], **TOL)
    np.testing.assert_allclose(f.P, G["xysr_ref_pred_P"], rtol=1e-12, atol=1e-9)
    f.update(z1)
    np.testing.assert_allclose(f.x, G["xysr_ref_upd_x"], rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(f.P, G["xysr_ref_upd_P"], rtol=1e-9, atol=1e-9)


def test_xysr_unfreeze_regression_scenario_of_the_reference():
    """tests/unit/test_kalman_filters_modes.py:164-193 (issue #2207): two observations, five missed frames, a third one ->
    freeze, virtual-trajectory replay from the pre-gap measurement, update.  State and covariance after the gap and after
    the re-observation against the reference filter."""
    from oracle.deepocsort import XYSRFilter
    from tests.golden.make_kalman_golden import UNFREEZE_OBS

    f = XYSRFilter(np.zeros((4, 1)))
    f.x, f.P = G["unfreeze_x0"].copy(), G["unfreeze_P0"].copy()
    f.Q, f.R = G["unfreeze_Q"].copy(), G["unfreeze_R"].copy()
    for obs in UNFREEZE_OBS[:2]:
        f.predict()
        f.update(obs)
    for _ in range(5):
        f.predict()
        f.update(None)
    np.testing.assert_allclose(f.x, G["unfreeze_gap_x"], rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(f.P, G["unfreeze_gap_P"], rtol=1e-11, atol=1e-8)
    f.predict()
    f.update(UNFREEZE_OBS[2])
    np.testing.assert_allclose(f.x, G["unfreeze_x"], rtol=1e-10, atol=1e-8)
    np.testing.assert_allclose(f.P, G["unfreeze_P"], rtol=1e-9, atol=1e-7)
