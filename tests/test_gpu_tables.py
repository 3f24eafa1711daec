"""Device table build (fscl_build_tables: asc-bias.c, sm-spline.c:316-523 as CUDA kernels) against
the tables of the compiled reference.  Not bit-exact by construction (device exp/log differ from
glibc in the last ulp); the stated bounds: spline values within 5e-11 * max(1, |value|) where the model probability
is above 1e-130 (log p > -300) and within 1e-9 relative in the far tail (towards denormals, where
the reference itself substitutes log(DBL_MIN) for p == 0); CLR within 1e-9 relative."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests.common import clr_close, problem

pytestmark = pytest.mark.gpu


def _ref_fsp(tab, **kw):
    rp = po.RefProblem(tab, **kw)
    flat = rp.flatten()
    fsp = [np.array([rp.fsp[d][f] for f in range(int(n) + 1)]) for d, n in enumerate(flat["depths"])]
    return flat, fsp


# K >= 2 is excluded: there the reference's tables hinge on ascprob_subsample(1, ...), which is pure
# rounding noise of either libm (D14 in DESIGN.md) -- there is nothing to agree on.
@pytest.mark.parametrize("kind,kw", [("small50", {}), ("asc1", dict(asc_depth=20, asc_min_freq=1)),
                                     ("asc1", dict(asc_depth=20, asc_min_freq=1, ascbias_background_only=1)),
                                     ("genome", {})])
def test_device_tables_match_reference_tables(built, kind, kw):
    from fscl_b200 import api
    tab, _, G = problem(kind)
    flat, fsp = _ref_fsp(tab, **kw)
    coef, null, flags = api.build_tables(flat["depths"], fsp, asc_depth=kw.get("asc_depth", 0),
                                         asc_min_freq=kw.get("asc_min_freq", 1),
                                         ascbias_background_only=kw.get("ascbias_background_only", 0))
    assert flags == 0
    assert np.array_equal(null, flat["null_tab"])                       # host libm, bit-exact
    used = np.unique(flat["cls"])
    L = po.lib()
    rng = np.random.default_rng(0)
    worst = worst_tail = 0.0
    for c in used[:: max(1, len(used) // 60)]:
        for x in rng.uniform(-20.0, 4.0, 40):
            a = L.fo_spline_eval(coef[c].ctypes.data, 200, flat["step"], float(x))
            b = L.fo_spline_eval(flat["coef"][c].ctypes.data, 200, flat["step"], float(x))
            if b > -300.0:
                worst = max(worst, abs(a - b) / max(1.0, abs(b)))
            else:
                worst_tail = max(worst_tail, abs(a - b) / abs(b))
    assert worst < 5e-11, worst
    assert worst_tail < 1e-9, worst_tail
    # the scan with device-built tables equals the scan with the reference's tables
    f2 = dict(flat)
    f2["coef"] = coef
    a = api.Scanner(flat).scan_chromosome(G)
    b = api.Scanner(f2).scan_chromosome(G)
    assert np.array_equal(a["sweep_pos"], b["sweep_pos"]) and np.array_equal(a["lalpha"], b["lalpha"])
    assert clr_close(a["clr"], b["clr"]).all()
