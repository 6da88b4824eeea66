"""Layout contract: flat moment order == sorted (n, l, m), m >= 0 (indexing.py:13-228, README.md:15, 91-97,
167-187).  The reference has no tests; these are its structural known answers + bijection properties."""
import numpy as np
import pytest

from zernike3d_b200 import indexing as ix


def test_readme_counts():
    # README.md:15 "order 50 ... ~12,051 Zernike polynomials ... 675 radial and 1326 spherical harmonics"
    assert ix.mu_(50) + 1 == 12051 == ix.num_moments(50)
    assert ix.mu_Y(50) == 1326
    assert ix.mu_R(50) == 675 and ix.num_radial(50) == 676  # 675 is the last 0-based index
    assert ix.num_moments(60) == 20336 and ix.num_radial(60) == 961 and ix.num_harmonics(60) == 1891


def test_readme_key_order():
    # README.md:179-187 display_moments(n_max=...) prints keys in this order
    want = [(0, 0, 0), (1, 1, 0), (1, 1, 1), (2, 0, 0), (2, 2, 0), (2, 2, 1), (2, 2, 2)]
    assert [ix.mu_to_nlm(i) for i in range(7)] == want
    # README.md:91-97: exactly entries 0,1,3,4,7,9 of arr_0[0:10] have zero imaginary part -> they are the m = 0 ones
    assert [i for i in range(10) if ix.mu_to_nlm(i)[2] == 0] == [0, 1, 3, 4, 7, 9]


@pytest.mark.parametrize("order", [0, 1, 2, 7, 50, 61])
def test_bijections(order):
    t = ix.nlm_table(order)
    keys = [tuple(int(v) for v in r) for r in t]
    assert keys == sorted(keys)
    assert all(0 <= m <= l <= n and (n - l) % 2 == 0 for n, l, m in keys)
    for mu in range(0, len(keys), max(1, len(keys) // 400)):
        assert ix.mu_to_nlm(mu) == keys[mu]
        assert ix.nlm_to_mu(keys[mu]) == mu
    n, l, m = keys[-1]
    assert ix.nlm_to_mu((n, l, -m)) == len(keys) - 1  # |m| is used, ix:181
    nl = ix.nl_table(order)
    for mu, (n, l, base) in enumerate(nl):
        assert ix.mu_R_to_nl(mu) == (n, l) and ix.nl_to_mu_R((n, l)) == mu
        assert keys[base] == (n, l, 0)
    for mu in range(ix.mu_Y(order)):
        assert ix.lm_to_mu_Y(ix.mu_Y_to_lm(mu)) == mu


def test_memory_predictor_matches_reference_figure():
    # ix:230-234 / memory_usage_Figure_1: 64^3 ~ 3.5 GB, 128^3 ~ 27.9 GB, 256^3 ~ 223 GB at order 50
    assert abs(ix.memory_requirement(50, 64) - 3.49) < 0.05
    assert abs(ix.memory_requirement(50, 128) - 27.9) < 0.1
    assert abs(ix.memory_requirement(50, 256) - 223) < 1


def test_top_level_module_like_reference():
    import indexing  # the reference is used as `from indexing import *`

    assert indexing.mu_to_nlm(12050) == (50, 50, 50)
    assert set(["mu_R", "nl_to_mu_R", "mu_R_to_nl", "mu_Y", "lm_to_mu_Y", "mu_Y_to_lm", "mu_", "nlm_to_mu",
                "mu_to_nlm"]) <= set(dir(indexing))
