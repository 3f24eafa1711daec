"""p-value projection (reference README.md:11, unimplemented upstream -> defined in DESIGN.md 3.6).
Host-only: non-central chi-square survival function vs scipy, and the moment estimate."""
import numpy as np
import pytest
from scipy.stats import chi2, ncx2


@pytest.fixture(scope="module")
def api(built):
    from fscl_b200 import api as a
    a.load_library()
    return a


def test_ncx2_survival_matches_scipy_into_the_far_tail(api):
    x = np.array([0.01, 0.5, 2.0, 7.0, 20.0, 60.0, 150.0, 400.0, 900.0])
    for df in (1.0, 2.0, 3.5):
        for lam in (0.0, 0.3, 2.0, 17.0, 250.0):
            got = api.ncx2_sf(x, df, lam)
            want = ncx2.sf(x, df, lam) if lam > 0 else chi2.sf(x, df)
            ok = want > 1e-290
            assert np.allclose(got[ok], want[ok], rtol=1e-10, atol=0)       # tolerance: 1e-10 relative
    assert (np.diff(api.ncx2_sf(np.linspace(0, 300, 200), 2.0, 5.0)) <= 0).all()


def test_projection_uses_moment_estimate_and_goes_below_one_over_n(api):
    rng = np.random.default_rng(0)
    n_pts, df = 6, 2.0
    pts = np.zeros(n_pts, dtype=api.PT_DTYPE)
    clr = np.zeros((n_pts, api.NULL_SAVE), dtype=np.float32)
    lam_true = [0.0, 1.0, 4.0, 9.0, 4.0, 4.0]
    n_perm = [1000, 1000, 1000, 10000, 25, 0]
    for i in range(n_pts):
        m = min(n_perm[i], api.NULL_SAVE)
        if m:
            clr[i, :m] = ncx2.rvs(df, lam_true[i], size=m, random_state=rng) if lam_true[i] > 0 else \
                chi2.rvs(df, size=m, random_state=rng)
        pts[i]["permute_n"] = n_perm[i]
        pts[i]["clr"] = [3.0, 10.0, 60.0, 120.0, 15.0, 7.0][i]
    lam, p = api.project_pvalues(pts, clr, df)
    for i in range(n_pts):
        m = min(n_perm[i], api.NULL_SAVE)
        if m == 0:
            assert p[i] == 1.0 and lam[i] == 0.0
            continue
        want_lam = max(float(clr[i, :m].astype(np.float64).mean()) - df, 0.0)
        assert abs(lam[i] - want_lam) <= 1e-12 * max(want_lam, 1.0)
        want_p = ncx2.sf(pts[i]["clr"], df, want_lam) if want_lam > 0 else chi2.sf(pts[i]["clr"], df)
        assert abs(p[i] - want_p) <= 1e-10 * want_p
    assert p[2] < 1.0 / n_perm[2] and p[3] < 1.0 / n_perm[3]       # the point of the projection
