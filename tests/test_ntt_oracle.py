"""CPU checks of the negacyclic-NTT restatement (oracle/ntt.hpp) and of the product's host-side root rule.

The reference has no transform of this shape, so the oracle is pinned by the definition itself:
the O(N^2) evaluation, the schoolbook product mod X^N + 1, and the stated root rule."""
import numpy as np
import pytest

import latticefold_b200 as lf
from latticefold_b200 import synth

P = {0: 0xFFFFFFFF00000001, 1: 2013265921}


def rnd(field, shape, seed):
    n = int(np.prod(shape))
    return synth.uniform_field(P[field], n, seed).reshape(shape)


@pytest.mark.parametrize("field", [0, 1])
def test_root_rule(oracle, field):
    p = P[field]
    for log_n in (4, 8, 10, 16):
        psi = oracle.ntt_root(field, log_n); n = 1 << log_n
        assert pow(psi, n, p) == p - 1                      # primitive 2N-th root
        assert psi == lf.ntt_root(field, log_n)             # the product's host code follows the same rule
        assert pow(oracle.ntt_root(field, log_n + 1), 2, p) == psi      # one tower
        if field == 0:
            assert pow(psi, n // 16, p) == 64               # radix-16 twiddles are powers of two


@pytest.mark.parametrize("field", [0, 1])
@pytest.mark.parametrize("log_n", [4, 5, 8])
def test_fast_matches_definition(oracle, field, log_n):
    a = rnd(field, (3, 1 << log_n), 11 + log_n)
    f = oracle.ntt(field, log_n, a); i = oracle.ntt(field, log_n, f, inverse=True)
    for r in range(3):
        assert np.array_equal(f[r], oracle.ntt_naive(field, log_n, a[r]))
        assert np.array_equal(oracle.ntt_naive(field, log_n, f[r], inverse=True), a[r])
    assert np.array_equal(i, a)


@pytest.mark.parametrize("field", [0, 1])
def test_convolution_theorem(oracle, field):
    log_n = 6; p = P[field]
    a, b = rnd(field, (1, 64), 3), rnd(field, (1, 64), 4)
    fa, fb = oracle.ntt(field, log_n, a), oracle.ntt(field, log_n, b)
    prod = np.array([[int(x) * int(y) % p for x, y in zip(fa[0], fb[0])]], dtype=np.uint64)
    assert np.array_equal(oracle.ntt(field, log_n, prod, inverse=True)[0], oracle.ntt_schoolbook(field, log_n, a[0], b[0]))


def test_known_small_case(oracle):
    # X mod (X^N + 1) evaluates to psi^(2k+1) at the k-th root: the transform of the monomial X lists the odd powers of psi
    for field in (0, 1):
        log_n = 4; p = P[field]; psi = oracle.ntt_root(field, log_n)
        a = np.zeros((1, 16), dtype=np.uint64); a[0, 1] = 1
        assert [int(v) for v in oracle.ntt(field, log_n, a)[0]] == [pow(psi, 2 * k + 1, p) for k in range(16)]
