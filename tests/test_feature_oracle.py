"""The feature-engineering oracle (oracle/feature_oracle.py) against the outputs of the reference's own
statements (tests/golden/feature_engineering.npz, made by tests/golden/make_feature_golden.py), and its
B-spline restatement against scipy (patsy is not installed: that block is unpinned)."""
import os

import numpy as np
import pytest

from oracle import feature_oracle as fo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "feature_engineering.npz")


@pytest.fixture(scope="module")
def fg():
    z = np.load(GOLDEN)
    fips = z["in_fips"]
    start = np.r_[0, np.flatnonzero(fips[1:] != fips[:-1]) + 1, len(fips)]
    return z, start


def test_feature_oracle_matches_the_reference_statements(fg):
    z, start = fg
    out = fo.engineer(start, z["in_date_days"], z["in_himax_c"], z["in_alert"], z["in_issued_in_advance"])
    for c in fo.INT_COLUMNS:
        np.testing.assert_array_equal(out[c], z["out_" + c].astype(np.int64), err_msg=c)
    exact = ("hi_max", "heat_qi", "heat_qi_above_25", "heat_qi_above_75", "hi_max_above_25", "hi_max_above_75",
             "hi_max*heat_qi", "hi_max_above_25*heat_qi", "hi_max_above_75*heat_qi", "alerts_2wks", "issued_in_advance")
    for c in fo.FLOAT_COLUMNS:
        if c in exact:  # no accumulation involved: the same IEEE operations in the same order
            np.testing.assert_array_equal(out[c], z["out_" + c].astype(np.float64), err_msg=c)
        else:  # rolling means: pandas' online Kahan sum vs a direct window sum
            np.testing.assert_allclose(out[c], z["out_" + c], rtol=1e-12, atol=1e-12, err_msg=c)
    # the synthetic frame exercises what it should: ties in the rank, ragged years, streaks, a year gap
    assert len(np.unique(z["out_heat_qi"])) < len(z["out_heat_qi"]) and z["out_alert_streak"].max() >= 4
    assert z["out_dos"].max() == 152 and (np.diff(start) % 153 != 0).any()
    # the date strings the reference writes (:182) are the civil dates of the day numbers
    d = z["in_date_days"].astype("datetime64[D]").astype(str)
    np.testing.assert_array_equal(d, z["out_date_str"])
    assert (fo.civil_year(z["in_date_days"]) == z["out_year"]).all()


def test_holiday_stub_of_the_golden_is_the_summer_federal_calendar(fg):
    """The golden's `holiday` column came from the stub calendar injected into the reference's statements (:184-187):
    Memorial Day, 4th of July (+ observed), Labor Day.  The device path takes this column as an input."""
    z, _ = fg
    days = z["in_date_days"][z["in_holiday"] == 1].astype("datetime64[D]")
    md = sorted({str(d)[5:] for d in days})
    assert "07-04" in md and all(m[:2] in ("05", "07", "09") for m in md)
    np.testing.assert_array_equal(z["in_holiday"], z["out_holiday"])


def test_bspline_restatement_matches_scipy():
    """`bs(x, df=3, degree=3, lower_bound=0, upper_bound=M+1) - 1` has no interior knots: the basis is the cubic
    B-spline basis on [0,0,0,0,M+1,...] minus its first function (patsy/splines.py) -- checked against scipy."""
    from scipy.interpolate import BSpline

    dos = np.tile(np.arange(153), 5)[:700]
    M = dos.max()
    x = dos / M
    knots = np.r_[[0.0] * 4, [float(M + 1)] * 4]
    design = BSpline.design_matrix(x, knots, 3).toarray()[:, 1:]
    np.testing.assert_allclose(fo.bernstein3(x / (M + 1)), design, rtol=1e-12, atol=1e-300)
    cols, mean, std = fo.bspline_dos(dos)
    ref = (design - design.mean(axis=0)) / design.std(axis=0, ddof=1)
    np.testing.assert_allclose(cols, ref, rtol=1e-9, atol=1e-12)
    assert np.allclose(cols.mean(axis=0), 0, atol=1e-9) and np.allclose(cols.std(axis=0, ddof=1), 1)


def test_holiday_flags_equal_the_golden_calendar(fg):
    """`weather2alert_b200.features.holiday_flags` (the product's stand-in for the `holidays` package, :184-187)
    reproduces the calendar the golden's holiday column was made with."""
    from weather2alert_b200.features import holiday_flags, us_summer_holidays

    z, _ = fg
    np.testing.assert_array_equal(holiday_flags(z["in_date_days"]), z["in_holiday"])
    import datetime as dt

    cal = us_summer_holidays([2015, 2021])
    assert dt.date(2015, 7, 3) in cal and dt.date(2015, 5, 25) in cal and dt.date(2015, 9, 7) in cal  # 4 July 2015 was a Saturday
    assert dt.date(2021, 6, 19) in cal and dt.date(2021, 7, 5) in cal
