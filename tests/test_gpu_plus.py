"""GPU parity tests of the LatticeFold+ consumers (latticefold_b200/csrc/lfplus.cu) against the CPU oracle (oracle/lfplus.hpp):
bit-exact proof images on the reference's own test cases (setchk.rs:358-495, rgchk.rs:344-433) and on seeded ones, both verifiers on
both provers' outputs, and size-independent properties at the reference's benchmark sizes (benches/utils/mod.rs)."""
import numpy as np
import pytest

import latticefold_b200 as lf
from latticefold_b200 import plus
from tests import plus_cases as pc

pytestmark = pytest.mark.gpu
RING = pc.RING_FROG


@pytest.fixture(scope="module")
def ctx():
    c = lf.Context(RING)
    yield c
    c.close()


def seeded(seed=None):
    t = plus.PoseidonTranscript()
    if seed is not None:
        t.absorb_base(np.asarray(seed, dtype=np.uint64))
    return t


@pytest.mark.parametrize("name", sorted(pc.set_check_cases()))
def test_set_check_bit_exact(ctx, oracle, name):
    nvars, sets, M, accept = pc.set_check_cases()[name]
    for seed in (None, [3, 1, 4]):
        got = plus.In(ctx, nvars, sets).set_check(M, seeded(seed))
        want = oracle.plus_set_check(RING, nvars, sets, M, seed=seed)
        assert got.shape == want.shape and np.array_equal(got, want)
        assert plus.set_check_verify(got, seeded(seed)) is accept
        assert oracle.plus_set_check_verify(RING, got, seed=seed) is accept


def test_set_check_prover_leaves_transcript_in_verifier_state(ctx):
    nvars, sets, M, _ = pc.set_check_cases()["rect_with_M"]
    tp, tv = seeded(), seeded()
    out = plus.In(ctx, nvars, sets).set_check(M, tp)
    assert plus.set_check_verify(out, tv)
    assert tp.get_challenge() == tv.get_challenge()


def test_set_check_rejects_bad_shapes(ctx):
    nvars, sets, M, _ = pc.set_check_cases()["rect_2sets"]
    with pytest.raises(lf.LfError):
        plus.In(ctx, 5, sets).set_check(M, seeded())      # 64 rows do not fit 2^5
    with pytest.raises(lf.LfError):
        plus.In(ctx, nvars, [sets[2]]).set_check(M, seeded())      # no matrix set (setchk.rs:85 panics)
    with pytest.raises(lf.LfError):
        plus.In(ctx, nvars, sets).set_check([pc.identity(32)], seeded())      # M of another width


@pytest.mark.parametrize("n,kappa,k,L,with_M", [(1 << 14, 1, 2, 1, False), (1 << 14, 1, 2, 1, True), (1 << 15, 2, 2, 2, True), (1 << 15, 1, 3, 1, False)])
def test_from_f_and_range_check_bit_exact(ctx, oracle, n, kappa, k, L, with_M):
    l, nvars = pc.frog_l(), n.bit_length() - 1
    fs, A = pc.range_check_inputs(n, kappa, seed=n + kappa + k, L=L, k=k)
    if L == 1 and k == 2 and not with_M:
        fs = pc.reference_range_check_f(n)      # the reference's own witness
    M = [pc.random_ring_sparse(n, n, 2, 77, constant=True)] if with_M else []
    Ad = plus.Matrix(ctx, A)
    inst = [plus.RgInstance.from_f(ctx, fs[i], Ad, 8, k, l) for i in range(L)]
    for i in range(L):
        tau, fc, cm = inst[i].read()
        otau, ofc, ocm = oracle.plus_rg_from_f(RING, fs[i], A, 8, k, l)
        assert np.array_equal(cm, ocm) and np.array_equal(tau, otau) and np.array_equal(fc, ofc)
    got = plus.Rg(ctx, nvars, inst).range_check(M, seeded([2, 7]))
    want = oracle.plus_range_check(RING, nvars, fs, A, 8, k, l, M, seed=[2, 7])
    assert got.shape == want.shape and np.array_equal(got, want)
    assert plus.range_check_verify(got, seeded([2, 7])) and oracle.plus_range_check_verify(RING, got, seed=[2, 7])
    assert not plus.range_check_verify(got, seeded())


def test_from_f_errors(ctx):
    n, kappa = 1 << 14, 1
    fs, A = pc.range_check_inputs(n, kappa, seed=1)
    Ad = plus.Matrix(ctx, A)
    big = fs[0].copy(); big[5, 3] = 1000      # needs more than two base-8 digits
    with pytest.raises(lf.LfError) as e:
        plus.RgInstance.from_f(ctx, big, Ad, 8, 2, pc.frog_l())
    assert e.value.code == -9
    with pytest.raises(lf.LfError):
        plus.RgInstance.from_f(ctx, fs[0][: n // 2], Ad, 8, 2, pc.frog_l())      # wrong witness length
    fs2, A2 = pc.range_check_inputs(1 << 12, 1, seed=2)
    with pytest.raises(lf.LfError):
        plus.RgInstance.from_f(ctx, fs2[0], plus.Matrix(ctx, A2), 8, 2, pc.frog_l())      # n below the length of tau (utils.rs:34-40)
    nc = fs[0].copy(); nc[0, 0] = np.uint64(pc.P_FROG)
    with pytest.raises(lf.LfError):
        plus.RgInstance.from_f(ctx, nc, Ad, 8, 2, pc.frog_l())


def test_range_check_at_benchmark_size(ctx):
    """benches/utils/mod.rs range_check::WITNESS_SCALING (n, k, kappa) = (2^17, 2, 2): too slow for the oracle; checked by
    acceptance, by linearity of the commitments and by the recomposition identity of tau."""
    n, kappa, k, l = 1 << 17, 2, 2, pc.frog_l()
    fs, A = pc.range_check_inputs(n, kappa, seed=99, L=1)
    Ad = plus.Matrix(ctx, A)
    inst = plus.RgInstance.from_f(ctx, fs[0], Ad, 8, k, l)
    tau, fc, cm = inst.read()
    st = np.array([int(t) - pc.P_FROG if int(t) > pc.P_FROG // 2 else int(t) for t in tau[: kappa * k * 16 * l * 16]], dtype=object).reshape(kappa * k * 16, l, 16)
    rec = [sum(int(st[0, i, c]) * 8 ** i for i in range(l)) % pc.P_FROG for c in range(16)]
    assert rec == [int(x) for x in cm[0, 0, 0]]
    f2 = fs[0].copy(); f2[:, :] = 0
    z = plus.RgInstance.from_f(ctx, f2, Ad, 8, k, l).read()[1]
    assert not z[0].any()      # cm_f of the zero witness
    dcom = plus.Rg(ctx, 17, [inst]).range_check([], seeded())
    assert plus.range_check_verify(dcom, seeded())
    t = dcom.copy(); t[40] = (int(t[40]) + 1) % pc.P_FROG
    assert not plus.range_check_verify(t, seeded())


@pytest.mark.parametrize("kappa,with_M,L", [(2, True, 1), (1, False, 1), (2, True, 2), (4, False, 1)])
def test_commitment_transformation_bit_exact(ctx, oracle, kappa, with_M, L):      # cm.rs:621-665 (test_com) and variations
    n, k, l, nvars = 1 << (16 if kappa == 4 else 15), 2, pc.frog_l(), 16 if kappa == 4 else 15
    fs, A = pc.range_check_inputs(n, kappa, seed=21 + kappa, L=L)
    if L == 1 and kappa == 2:
        fs = pc.reference_range_check_f(n)
    M = []
    if with_M:
        m = pc.identity(n); m["val"] = m["val"].copy(); m["val"][0, 0] = 2
        M = [m] if L == 1 else [m, pc.random_ring_sparse(n, n, 2, 31, constant=True)]
    Ad = plus.Matrix(ctx, A)
    inst = [plus.RgInstance.from_f(ctx, fs[i], Ad, 8, k, l) for i in range(L)]
    tp = seeded([8, 1])
    proof, comx, g = plus.Cm(plus.Rg(ctx, nvars, inst)).prove(M, tp)
    oproof, ocomx, og = oracle.plus_cm_prove(RING, nvars, fs, A, 8, k, l, M, seed=[8, 1])
    assert proof.shape == oproof.shape and np.array_equal(proof, oproof)
    assert np.array_equal(comx, ocomx) and np.array_equal(g, og)
    tv = seeded([8, 1])
    ok, comx_v = plus.cm_verify(proof, len(M), tv, nvars=nvars, L=L, kappa=kappa)
    assert ok and np.array_equal(comx_v, comx)
    assert tp.get_challenge() == tv.get_challenge()      # prover and verifier leave the transcript in the same state
    assert oracle.plus_cm_verify(RING, proof, M, seed=[8, 1])[0]
    cmg = comx[: L * kappa * 16].reshape(L, kappa, 16)
    for li in range(L):
        assert np.array_equal(oracle.plus_mat_vec(RING, A, g[li]), cmg[li])      # A g = cm_g: g opens the folded commitment


def test_commitment_transformation_at_benchmark_size(ctx, oracle):
    """benches/utils/mod.rs commitment_transform rows use n = 2^16, k = 2, kappa = 2: acceptance by both verifiers and A g = cm_g."""
    n, kappa, k, l, nvars = 1 << 16, 2, 2, pc.frog_l(), 16
    fs, A = pc.range_check_inputs(n, kappa, seed=5, L=1)
    inst = [plus.RgInstance.from_f(ctx, fs[0], plus.Matrix(ctx, A), 8, k, l)]
    proof, comx, g = plus.Cm(plus.Rg(ctx, nvars, inst)).prove([], seeded())
    assert plus.cm_verify(proof, 0, seeded(), nvars=nvars, L=1, kappa=kappa)[0]
    assert oracle.plus_cm_verify(RING, proof, [])[0]
    assert np.array_equal(oracle.plus_mat_vec(RING, A, g[0]), comx[: kappa * 16].reshape(kappa, 16))


@pytest.mark.parametrize("L,n_M", [(2, 2), (3, 0)])
def test_mlin_and_decompose_bit_exact(ctx, oracle, L, n_M):      # mlin.rs:41-106, decomp.rs:32-127 (the data flow of test_decomp_g from the folded instances on)
    n, kappa, k, l = 1 << 15, 2, 2, pc.frog_l()
    fs, A = pc.range_check_inputs(n, kappa, seed=50 + L, L=L)
    m = pc.identity(n); m["val"] = m["val"].copy(); m["val"][0, 0] = 2
    M = [m, pc.random_ring_sparse(n, n, 2, 51, constant=True)][:n_M]
    Ad = plus.Matrix(ctx, A)
    proof, x, g = plus.Mlin(ctx, fs, 8, k, l).mlin(Ad, M, seeded([3]))
    oproof, ox, og = oracle.plus_mlin(RING, fs, A, 8, k, l, M, seed=[3])
    assert np.array_equal(proof, oproof) and np.array_equal(g, og) and all(np.array_equal(x[key], ox[key]) for key in ("cm_g", "ro", "vo"))
    assert plus.cm_verify(proof, len(M), seeded([3]))[0]
    for B in (int(pc.P_FROG ** 0.5) + 2, 1 << 12):      # decomp.rs:190-193, and a base that makes both digits non-trivial
        dproof, F = plus.decompose(ctx, Ad, g, x["ro"], M, B)
        odproof, oF = oracle.plus_decompose(RING, g, x["ro"], A, B, M)
        assert np.array_equal(dproof, odproof) and np.array_equal(F, oF)
        assert plus.decompose_verify(dproof, kappa, len(M), x["cm_g"], x["vo"], B) and oracle.plus_decompose_verify(RING, dproof, kappa, len(M), x["cm_g"], x["vo"], B)
    with pytest.raises(lf.LfError) as e:      # two digits in base 4 cannot hold g
        plus.decompose(ctx, Ad, g, x["ro"], M, 4)
    assert e.value.code == -9


@pytest.mark.parametrize("n", [1 << 7, 1 << 12, 96])
def test_r1cs_linearize_bit_exact(ctx, oracle, n):      # r1cs.rs:207-232; n = 96: a witness shorter than 2^nvars
    abc, f = pc.r1cs_instance(n, n + 1)
    tp = seeded([4])
    linb, lp = plus.ComR1CS(ctx, abc, f).linearize(tp)
    to = oracle.plus_transcript(RING, [4])
    olinb, olp = oracle.plus_r1cs_linearize(RING, abc, f, to)
    assert np.array_equal(lp, olp) and np.array_equal(linb["r"], olinb["r"]) and np.array_equal(linb["v"], olinb["v"])
    assert plus.r1cs_linearize_verify(lp, seeded([4])) and oracle.plus_r1cs_linearize_verify(RING, lp, oracle.plus_transcript(RING, [4]))
    assert tp.get_challenge() == oracle.plus_transcript_challenge(to)
    # general ring-valued matrices and witness (no relation holds: the proof must be rejected, but the images still agree)
    rng = np.random.default_rng(n)
    g = rng.integers(0, pc.P_FROG, size=f.shape, dtype=np.uint64)
    abc2 = [pc.random_ring_sparse(n, n, 2, s) for s in (1, 2, 3)]
    _, lp2 = plus.ComR1CS(ctx, abc2, g).linearize(seeded())
    _, olp2 = oracle.plus_r1cs_linearize(RING, abc2, g, oracle.plus_transcript(RING))
    assert np.array_equal(lp2, olp2) and not plus.r1cs_linearize_verify(lp2, seeded())


def test_plus_prover_three_folds_bit_exact(ctx, oracle):      # plus.rs:216-272 (test_prove_multi: k = 4, n = 2^16, three folds of the accumulator with a fresh instance)
    n, kappa, k, l, B = 1 << 16, 2, 4, pc.frog_l(), 3000
    _, A = pc.range_check_inputs(n, kappa, seed=61)
    abc, f0 = pc.r1cs_instance(n, 5)
    oflow = pc.OraclePlus(oracle, A, abc, 8, k, l, B)
    prover = plus.PlusProver(ctx, plus.Matrix(ctx, A), abc, 8, k, l, B, seeded())
    verifier = plus.PlusVerifier(kappa, len(abc), B, seeded())
    for fold in range(3):
        proof = prover.prove([plus.ComR1CS(ctx, abc, f0)])
        want = oflow.prove([(abc, f0)])
        assert np.array_equal(proof["cmproof"], want["cmproof"]) and np.array_equal(proof["dproof"], want["dproof"]), fold
        assert all(np.array_equal(a, b) for a, b in zip(proof["lproof"], want["lproof"]))
        assert all(np.array_equal(proof["linb2x"][key], want["linb2x"][key]) for key in ("cm_g", "ro", "vo"))
        assert all(np.array_equal(a, b) for a, b in zip(prover.acc_download(), oflow.acc))      # the accumulated witnesses (device resident between the folds)
        assert verifier.verify(proof) and oflow.verify(proof)


def test_plus_prove_reference_setup(ctx, oracle):      # plus.rs:161-214 (test_prove) with the reference's own construction of the instances
    d, n, k, kappa, L = 16, 1 << 15, 2, 2, 3
    B = plus.estimate_bound(d * 128, L, d, k) + 1      # utils.rs:105-115
    assert B == 6186
    m, l = n // k, pc.frog_l()
    rng = np.random.default_rng(9)
    zs = [np.zeros((m, d), dtype=np.uint64) for _ in range(2)]
    for z in zs:
        z[:, 0] = rng.integers(0, 2, size=m).astype(np.uint64)
    abc = plus.r1cs_decomposed_square([pc.identity(m)] * 3, n, B, k)      # r1cs.rs:171-185
    assert abc[0]["nrows"] == n and abc[0]["ncols"] == n
    comps = [plus.com_r1cs_new(ctx, abc, z, B, k) for z in zs]
    for c_, z in zip(comps, zs):
        assert np.array_equal(c_.f, oracle.gadget_decompose(RING, z, B, k))      # ComR1CS::new's f
        # (A f) o (B f) = C f row by row: the decomposed system holds the relation of the original one (check_relation)
        zf = np.zeros((n, d), dtype=np.uint64); zf[:m] = z      # identity rows recompose z
        assert np.array_equal(c_.f[0::k][:, 0], z[:, 0])
    _, A = pc.range_check_inputs(n, kappa, seed=3)
    prover = plus.PlusProver(ctx, plus.Matrix(ctx, A), abc, 8, k, l, B, seeded())
    proof = prover.prove(comps)
    assert plus.PlusVerifier(kappa, 3, B, seeded()).verify(proof)
    oflow = pc.OraclePlus(oracle, A, abc, 8, k, l, B)
    want = oflow.prove([(abc, c_.f) for c_ in comps])
    assert np.array_equal(proof["cmproof"], want["cmproof"]) and np.array_equal(proof["dproof"], want["dproof"]) and oflow.verify(proof)
    t = dict(proof); t["lproof"] = [proof["lproof"][1], proof["lproof"][0]]
    assert not plus.PlusVerifier(kappa, 3, B, seeded()).verify(t)
