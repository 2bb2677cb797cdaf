"""CPU tests of the LatticeFold+ oracle (oracle/lfplus.hpp): the reference's own set-check / range-check tests
(crates/latticefold-plus/src/setchk.rs:358-495, rgchk.rs:344-433: prove -> verify accepts, non-monomials are rejected)
and the identities that pin exp / psi / split / tensor."""
import numpy as np
import pytest

from tests import plus_cases as pc

RING = pc.RING_FROG


@pytest.mark.parametrize("name", sorted(pc.set_check_cases()))
def test_set_check_cases(oracle, name):
    nvars, sets, M, accept = pc.set_check_cases()[name]
    out = oracle.plus_set_check(RING, nvars, sets, M)
    assert oracle.plus_set_check_verify(RING, out) is accept
    if accept:      # any change to the proof image is rejected (sumcheck messages, evaluations)
        for pos in (5 + nvars + 3, 5 + nvars + nvars * 64 + 1):      # a sumcheck message, an evaluation e[0] (Out.r and e[1..] are not read by this verifier)
            t = out.copy()
            t[pos] = (int(t[pos]) + 1) % pc.P_FROG
            assert not oracle.plus_set_check_verify(RING, t)
        assert not oracle.plus_set_check_verify(RING, out, seed=[1, 2, 3])      # another transcript state


def test_set_check_transcript_seed(oracle):
    nvars, sets, M, _ = pc.set_check_cases()["rect_2sets"]
    a = oracle.plus_set_check(RING, nvars, sets, M, seed=[7, 8, 9])
    b = oracle.plus_set_check(RING, nvars, sets, M)
    assert oracle.plus_set_check_verify(RING, a, seed=[7, 8, 9]) and not np.array_equal(a, b)


def test_ring_mul_negacyclic(oracle):
    rng = np.random.default_rng(1)
    a, b = (rng.integers(0, pc.P_FROG, size=16, dtype=np.uint64) for _ in range(2))
    ab = oracle.plus_ring_mul(RING, a, b)
    assert np.array_equal(ab, oracle.coeff_mul(RING, a, b))      # the schoolbook product of ring.hpp
    # X^15 * X = -1; monomial and constant operands take the rotation path
    m = oracle.plus_ring_mul(RING, pc.monomial(15), pc.monomial(1))
    assert int(m[0]) == pc.P_FROG - 1 and not m[1:].any()
    assert np.array_equal(oracle.plus_ring_mul(RING, a, pc.monomial(3)), oracle.coeff_mul(RING, a, pc.monomial(3)))


def test_tensor(oracle):      # utils.rs:118-131
    p = pc.P_FROG
    t = oracle.plus_tensor(RING, [10, 2])
    want = [(-9 * -1) % p, (-9 * 2) % p, (-10) % p, 20]
    assert [int(x) for x in t] == want


def test_from_f_double_commitment(oracle):
    """RgInstance::from_f: tau is the gadget decomposition of A * exp(D_f); the commitments are linear in their inputs."""
    n, kappa, k = 1 << 14, 1, 2
    fs, A = pc.range_check_inputs(n, kappa, seed=3)
    l = pc.frog_l()
    tau, fcoms, comM = oracle.plus_rg_from_f(RING, fs[0], A, 8, k, l)
    signed = np.array([int(t) - pc.P_FROG if int(t) > pc.P_FROG // 2 else int(t) for t in tau], dtype=np.int64)
    assert np.abs(signed).max() <= 4 and not tau[kappa * k * 16 * l * 16:].any()
    # recomposing the first entry's digits gives back comM_f[0][0][0]
    digs = signed[: l * 16].reshape(l, 16)
    rec = [sum(int(digs[i, c]) * 8 ** i for i in range(l)) % pc.P_FROG for c in range(16)]
    assert rec == [int(x) for x in comM[0, 0, 0]]
    # cm_f = A f as a ring product (checked on the first witness entries through linearity: commit of f with one entry)
    e0 = np.zeros_like(fs[0]); e0[0] = fs[0][0]
    _, fc0, _ = oracle.plus_rg_from_f(RING, e0, A, 8, k, l)
    assert np.array_equal(fc0[0, 0], oracle.plus_ring_mul(RING, A[0, 0], fs[0][0]))


@pytest.mark.parametrize("with_M", [False, True])
def test_range_check_reference_case(oracle, with_M):      # rgchk.rs:352-432 at n = 2^14 (2^15 in the reference; tau needs n > 11264)
    n, kappa, k, l = 1 << 14, 1, 2, pc.frog_l()
    f = pc.reference_range_check_f(n)
    _, A = pc.range_check_inputs(n, kappa, seed=5)
    M = []
    if with_M:      # test_range_check_mm: identity with entry (0, 0) = 2
        m = pc.identity(n); m["val"] = m["val"].copy(); m["val"][0, 0] = 2; M = [m]
    dcom = oracle.plus_range_check(RING, 14, f, A, 8, k, l, M)
    assert oracle.plus_range_check_verify(RING, dcom)
    h = [int(x) for x in dcom[5:10]]      # nvars, n_mat, ncols, n_vec, n_M
    v0 = 5 + 5 + h[0] + h[0] * 64 + (1 + h[4]) * h[1] * h[2] * 16 + h[3] * 16      # first word after the set-check image: v[0]
    t = dcom.copy(); t[v0] = (int(t[v0]) + 1) % pc.P_FROG
    assert not oracle.plus_range_check_verify(RING, t)
    t = dcom.copy(); t[-1] = (int(t[-1]) + 1) % pc.P_FROG      # the commitments are carried, not checked (rgchk.rs:190-246)
    assert oracle.plus_range_check_verify(RING, t)
    if with_M:      # c[1] = MLE(M f)(r) is held to the psi check of e[1]
        pos = dcom.size - 3 * kappa * 16 - 1
        t = dcom.copy(); t[pos] = (int(t[pos]) + 1) % pc.P_FROG
        assert not oracle.plus_range_check_verify(RING, t)


def test_range_check_random_two_instances(oracle):
    n, kappa, k, l = 1 << 15, 2, 2, pc.frog_l()
    fs, A = pc.range_check_inputs(n, kappa, seed=11, L=2)
    dcom = oracle.plus_range_check(RING, 15, fs, A, 8, k, l, [pc.random_ring_sparse(n, n, 2, 12, constant=True)])
    assert oracle.plus_range_check_verify(RING, dcom)
    bad = oracle.plus_range_check(RING, 15, fs, A, 8, k, l, [pc.random_ring_sparse(n, n, 2, 12)])      # ring-valued M: the psi tests of e[1] / c[1] fail
    assert not oracle.plus_range_check_verify(RING, bad)


@pytest.mark.parametrize("kappa,with_M,L", [(2, True, 1), (1, False, 1), (2, True, 2)])
def test_commitment_transformation(oracle, kappa, with_M, L):      # cm.rs:621-665 test_com (kappa = 2, M = identity with a 2) and variations
    n, k, l, nvars = 1 << 15, 2, pc.frog_l(), 15
    fs, A = pc.range_check_inputs(n, kappa, seed=21 + kappa, L=L)
    if L == 1:
        fs = pc.reference_range_check_f(n)
    M = []
    if with_M:
        m = pc.identity(n); m["val"] = m["val"].copy(); m["val"][0, 0] = 2; M = [m]
    proof, comx, g = oracle.plus_cm_prove(RING, nvars, fs, A, 8, k, l, M)
    ok, comx_v = oracle.plus_cm_verify(RING, proof, M, nvars=nvars, L=L, kappa=kappa)
    assert ok and np.array_equal(comx, comx_v)      # prover and verifier derive the same ComX
    for pos in (proof.size - 1, proof.size - 2 * L * (1 + len(M)) * 4 * 16 - 5):      # an evaluation of the second sumcheck, a message of it
        t = proof.copy(); t[pos] = (int(t[pos]) + 1) % pc.P_FROG
        assert not oracle.plus_cm_verify(RING, t, M)[0]
    assert not oracle.plus_cm_verify(RING, proof, M, seed=[5])[0]
    # g = s0 tau + s1 m_tau + s2 f + h is a commitment opening: A g = cm_g (the folded commitment of ComX), by linearity of A
    cmg = comx[: L * kappa * 16].reshape(L, kappa, 16)
    for li in range(L):
        assert np.array_equal(oracle.plus_mat_vec(RING, A, g[li]), cmg[li])


def mlin_inputs(n, kappa, L, seed):
    fs, A = pc.range_check_inputs(n, kappa, seed=seed, L=L)
    m = pc.identity(n); m["val"] = m["val"].copy(); m["val"][0, 0] = 2
    return fs, A, [m, pc.random_ring_sparse(n, n, 2, seed + 1, constant=True)]


def test_mlin_then_decompose(oracle):      # the data flow of decomp.rs:189-268 (test_decomp_g) from the folded instances on: mlin -> CmProof::verify -> decompose -> DecompProof::verify
    n, kappa, k, l, L = 1 << 15, 2, 2, pc.frog_l(), 2
    fs, A, M = mlin_inputs(n, kappa, L, 41)
    proof, x, g = oracle.plus_mlin(RING, fs, A, 8, k, l, M)
    ok, comx = oracle.plus_cm_verify(RING, proof, M, nvars=15, L=L, kappa=kappa)
    assert ok
    cmg = comx[: L * kappa * 16].reshape(L, kappa, 16)      # LinB2X sums the per-instance ComX
    assert [[int(v) for v in row] for row in x["cm_g"]] == [[sum(int(cmg[li, r, c]) for li in range(L)) % pc.P_FROG for c in range(16)] for r in range(kappa)]
    assert np.array_equal(oracle.plus_mat_vec(RING, A, g), x["cm_g"])
    B = int(pc.P_FROG ** 0.5) + 2      # decomp.rs:190-193
    for Bv in (B, 1 << 12):
        dproof, F = oracle.plus_decompose(RING, g, x["ro"], A, Bv, M)
        assert oracle.plus_decompose_verify(RING, dproof, kappa, len(M), x["cm_g"], x["vo"], Bv)
        t = dproof.copy(); t[3] = (int(t[3]) + 1) % pc.P_FROG
        assert not oracle.plus_decompose_verify(RING, t, kappa, len(M), x["cm_g"], x["vo"], Bv)
        t = dproof.copy(); t[-1] = (int(t[-1]) + 1) % pc.P_FROG
        assert not oracle.plus_decompose_verify(RING, t, kappa, len(M), x["cm_g"], x["vo"], Bv)
        # F0 + B F1 = g, digits inside the balanced range
        rec = (F[0].astype(object) + Bv * F[1].astype(object)) % pc.P_FROG
        assert np.array_equal(rec.astype(np.uint64), g)


def test_r1cs_linearize(oracle):      # r1cs.rs:207-232 (test_linearization): prove -> verify; an unsatisfied system is rejected
    n = 1 << 7
    abc, f = pc.r1cs_instance(n, 3)
    tp, tv = oracle.plus_transcript(RING), oracle.plus_transcript(RING)
    linb, lp = oracle.plus_r1cs_linearize(RING, abc, f, tp)
    assert oracle.plus_r1cs_linearize_verify(RING, lp, tv)
    assert oracle.plus_transcript_challenge(tp) == oracle.plus_transcript_challenge(tv)
    bad = f.copy(); bad[0, 0] = 3      # (2 * 3) * 3 != 2 * 3
    _, lp2 = oracle.plus_r1cs_linearize(RING, abc, bad, oracle.plus_transcript(RING))
    assert not oracle.plus_r1cs_linearize_verify(RING, lp2, oracle.plus_transcript(RING))


def test_plus_prover_fold_two_instances(oracle):      # plus.rs:161-214 (test_prove): two committed R1CS instances folded into one accumulator
    n, kappa, k, l = 1 << 15, 2, 2, pc.frog_l()
    _, A = pc.range_check_inputs(n, kappa, seed=61)
    abc, f0 = pc.r1cs_instance(n, 5)
    f1 = f0.copy(); f1[1:, 0] = 1 - f1[1:, 0]
    flow = pc.OraclePlus(oracle, A, abc, 8, k, l, 1 << 11)
    p1 = flow.prove([(abc, f0), (abc, f1)])
    assert flow.verify(p1)
    t = dict(p1); t["dproof"] = p1["dproof"].copy(); t["dproof"][0] ^= np.uint64(1)
    flow2 = pc.OraclePlus(oracle, A, abc, 8, k, l, 1 << 11)
    flow2.prove([(abc, f0), (abc, f1)])
    assert not flow2.verify(t)
