"""Pins the CPU oracle against every known-answer test the reference holds for the hot path,
plus the CRT-table-independent algebraic identities of 8c.  CPU only."""
import json
import os

import numpy as np
import pytest

from latticefold_b200 import synth
from tests.helpers import rand_elems

GOLD = os.path.join(os.path.dirname(__file__), "golden")
G, BB, FROG = synth.RING_GOLDILOCKS, synth.RING_BABYBEAR, synth.RING_FROG


def load(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def test_ring_shapes(oracle):
    # crates/cyclotomic-rings/src/rings/{goldilocks,babybear,frog}.rs:9-20
    for ring, R in synth.RINGS.items():
        i = oracle.info(ring)
        assert (i["p"], i["d"], i["S"], i["tau"]) == (R["p"], R["d"], R["S"], R["tau"])


def test_rot_lin_combination_kat(oracle):
    # crates/cyclotomic-rings/src/rotation.rs:174-776
    g = load("rotsum_goldilocks.json")
    rho = np.array(g["rho"], dtype=np.uint64).reshape(3, 24)
    theta = np.array(g["theta"], dtype=np.uint64).reshape(3, 3, 24)
    exp = np.array(g["expected"], dtype=np.uint64).reshape(3, 24)
    got = oracle.rot_lin_combination(G, rho, theta)
    assert np.array_equal(got, exp)


def test_transcript_kats(oracle):
    # crates/latticefold/src/transcript/poseidon.rs:86-142
    g = load("transcript_goldilocks.json")
    t = oracle.transcript(G)
    t.absorb_base(np.array(g["absorbed"], dtype=np.uint64))
    assert list(map(int, t.get_challenge())) == g["big_challenge"]
    t = oracle.transcript(G)
    t.absorb_base(np.array(g["absorbed"], dtype=np.uint64))
    assert list(map(int, t.get_short_challenge())) == g["small_challenge_coeffs"]


@pytest.mark.parametrize("ring,name", [(G, "goldilocks"), (BB, "babybear"), (FROG, "frog")])
def test_small_challenge_from_bytes(oracle, ring, name):
    # crates/cyclotomic-rings/src/rings/{goldilocks.rs:78-115,babybear.rs:78-114,frog.rs:66-95}
    g = load("challenge_sets.json")[name]
    got = oracle.short_challenge_from_bytes(ring, g["bytes"])
    exp = np.zeros(synth.RINGS[ring]["d"], dtype=np.uint64)
    exp[:len(g["coeffs"])] = np.array(g["coeffs"], dtype=np.uint64)
    assert np.array_equal(got, exp)


def test_commit_ntt_closed_form(oracle):
    # crates/latticefold/src/commitment/commitment_scheme.rs:142-160: A[i][j] = i*n + j, f = 2, kappa = 9, n = 2^15
    n, kappa, p = 1 << 15, 9, synth.RINGS[G]["p"]
    A = np.zeros((kappa, n, 24), dtype=np.uint64)
    vals = (np.arange(kappa, dtype=np.uint64)[:, None] * np.uint64(n) + np.arange(n, dtype=np.uint64)[None, :])
    A[:, :, ::3] = vals[:, :, None]
    f = np.zeros((n, 24), dtype=np.uint64)
    f[:, ::3] = 2
    cm = oracle.commit(G, A, f)
    for i in range(kappa):
        exp = np.zeros(24, dtype=np.uint64)
        exp[::3] = (n * (2 * i * n + (n - 1))) % p
        assert np.array_equal(cm[i], exp)


def test_commit_wrong_length(oracle):
    # commitment_scheme.rs:37-44 -> CommitmentError::WrongWitnessLength
    from oracle.pyoracle import OracleError
    A = rand_elems(G, 2 * 4, 1).reshape(2, 4, 24)
    with pytest.raises(OracleError) as e:
        oracle.commit(G, A, rand_elems(G, 3, 2))
    assert e.value.code == -1


def test_get_fhat_kat(oracle):
    # crates/latticefold/src/arith.rs:456-502
    f = np.zeros((2, 24), dtype=np.uint64)
    f[0, :3] = [1, 2, 3]
    f[1, :3] = [4, 5, 6]
    f[1, 3:] = 1
    fh, lens = oracle.fhat(G, f)
    def ntt(slots):
        e = np.zeros(24, dtype=np.uint64); e[::3] = slots; return e
    assert np.array_equal(fh[0, 0], ntt([1, 2, 3, 0, 0, 0, 0, 0]))
    assert np.array_equal(fh[0, 1], ntt([4, 5, 6, 1, 1, 1, 1, 1]))
    for j in (1, 2):
        assert np.array_equal(fh[j, 0], ntt([0] * 8)) and np.array_equal(fh[j, 1], ntt([1] * 8))
    assert list(lens) == [2, 2, 2]


def test_mat_vec_mul_kat(oracle):
    # crates/latticefold/src/arith/utils.rs:134-155 (embedded as constant ring elements)
    def c(v):
        e = np.zeros(24, dtype=np.uint64); e[::3] = v; return e
    M = dict(nrows=3, ncols=3, row_ptr=np.array([0, 1, 3, 4], dtype=np.uint64), col=np.array([0, 1, 2, 2], dtype=np.uint64),
             val=np.stack([c(1), c(2), c(1), c(3)]))
    z = np.stack([c(1), c(1), c(1)])
    assert np.array_equal(oracle.spmv(G, M, z), np.stack([c(1), c(3), c(3)]))
    from oracle.pyoracle import OracleError
    with pytest.raises(OracleError) as e:
        oracle.spmv(G, M, z[:2])
    assert e.value.code == -2


@pytest.mark.parametrize("ring", [G, BB, FROG])
def test_crt_identities(oracle, ring):
    # SURVEY 8(c) acceptance (i): ICRT(CRT a (.) CRT b) == schoolbook a*b mod Phi ; ICRT o CRT = id
    a, b = rand_elems(ring, 6, 11), rand_elems(ring, 6, 12)
    A, Bn = oracle.crt(ring, a), oracle.crt(ring, b)
    assert np.array_equal(oracle.icrt(ring, A), a)
    prod = oracle.icrt(ring, oracle.ntt_mul(ring, A, Bn))
    for i in range(6):
        assert np.array_equal(prod[i], oracle.coeff_mul(ring, a[i], b[i]))


@pytest.mark.parametrize("ring", [G, BB, FROG])
def test_rotsum_is_ring_product(oracle, ring):
    # crates/cyclotomic-rings/src/rotation.rs:115-135: RotSum(a, coeffs(b)) == coeffs(a*b)
    R = synth.RINGS[ring]
    a, b = rand_elems(ring, 1, 21), rand_elems(ring, 1, 22)
    # theta = b's coefficients embedded one per slot value: tau NTT elements whose flattened slot list is coeffs(b)
    theta = np.zeros((1, R["tau"], R["d"]), dtype=np.uint64)
    flat = theta.reshape(R["d"], R["tau"])
    flat[:, 0] = b[0]
    got = oracle.rot_lin_combination(ring, a, theta).reshape(R["d"], R["tau"])
    assert np.array_equal(got[:, 0], oracle.coeff_mul(ring, a[0], b[0]))
    assert not got[:, 1:].any()


@pytest.mark.parametrize("ring,B,L,b,K", [(G, 1 << 15, 5, 2, 15), (BB, 1 << 8, 4, 2, 8), (FROG, 1 << 8, 8, 2, 10), (G, 1 << 16, 4, 4, 8)])
def test_decompose_recompose(oracle, ring, B, L, b, K):
    # decomposition_parameters.rs:49-113 ; decomposition/utils.rs:84-195 (recompose o decompose = id, digit bounds)
    R = synth.RINGS[ring]; p = R["p"]
    a = rand_elems(ring, 5, 31)
    dec = oracle.gadget_decompose(ring, a, B, L)
    assert dec.shape == (5 * L, R["d"])
    signed = np.where(dec > p // 2, dec.astype(object) - p, dec.astype(object))
    assert (abs(signed) <= B // 2).all()
    assert np.array_equal(oracle.gadget_recompose(ring, dec, B, L), a)
    pieces = oracle.decompose_to_vec(ring, dec, b, K)
    sp = np.where(pieces > p // 2, pieces.astype(object) - p, pieces.astype(object))
    assert (abs(sp) <= b // 2).all()
    back = sum(sp[k] * (b ** k) for k in range(K)) % p
    assert np.array_equal(back.astype(np.uint64), dec)


def test_eq_table_matches_eq_eval(oracle):
    # sumcheck/utils.rs:78-170: table entry x == eq_eval(bits(x), r), r[0] on bit 0
    s = 4
    r = rand_elems(G, s, 41)
    tab = oracle.eq_table(G, r)
    for x in (0, 1, 6, 15):
        bits = np.zeros((s, 24), dtype=np.uint64)
        for i in range(s):
            if (x >> i) & 1:
                bits[i, ::3] = 1
        assert np.array_equal(tab[x], oracle.eq_eval(G, bits, r))


def test_evaluate_mles_is_eq_inner_product(oracle):
    # mle_helpers.rs:65-88: mle(r) = sum_x eq(x, r) mle[x]; truncated tail = zeros; wrong point length -> error
    from oracle.pyoracle import OracleError
    s = 3
    r = rand_elems(G, s, 51)
    m = rand_elems(G, 2 * 5, 52).reshape(2, 5, 24)
    got = oracle.evaluate_mles(G, m, s, r)
    tab = oracle.eq_table(G, r)
    p = synth.RINGS[G]["p"]
    for k in range(2):
        acc = np.zeros(24, dtype=object)
        for x in range(5):
            acc = (acc + oracle.ntt_mul(G, tab[x:x + 1], m[k, x:x + 1])[0].astype(object)) % p
        assert np.array_equal(got[k], acc.astype(np.uint64))
    with pytest.raises(OracleError):
        oracle.evaluate_mles(G, m, s, r[:2])
