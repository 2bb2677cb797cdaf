"""Parity of the CUDA path (through the C ABI) against the CPU oracle: bit-exact on identical seeded inputs.
Sizes are chosen so the oracle finishes in seconds; full-size properties are in test_gpu_fullsize.py."""
import numpy as np
import pytest

from latticefold_b200 import synth
from tests.helpers import rand_elems, rand_sf_broadcast

pytestmark = pytest.mark.gpu
GOLD, BB, FROG = synth.RING_GOLDILOCKS, synth.RING_BABYBEAR, synth.RING_FROG


@pytest.fixture(scope="module", params=[GOLD, BB, FROG], ids=["goldilocks", "babybear", "frog"])
def ctx(gpu, request):
    """One context per reference ring; every parity test below runs on all three (module-level G / P / D follow it)."""
    global G, P, D, TAU
    G = request.param
    P, D, TAU = synth.RINGS[G]["p"], synth.RINGS[G]["d"], synth.RINGS[G]["tau"]
    c = gpu.Context(G, 0)
    yield c
    c.close()


def const_el(v):
    e = np.zeros(D, dtype=np.uint64); e[::TAU] = v % P; return e


@pytest.mark.parametrize("n", [0, 1, 63, 64, 65, 1000])
def test_upload_download_roundtrip(ctx, n):
    a = rand_elems(G, n, 1)
    assert np.array_equal(ctx.upload(a).download(), a)


@pytest.mark.parametrize("n", [1, 127, 128, 1500])
def test_crt_icrt(ctx, oracle, n):
    a = rand_elems(G, n, 2)
    ntt = ctx.crt(ctx.upload(a, 1))
    assert np.array_equal(ntt.download(), oracle.crt(G, a))
    assert np.array_equal(ctx.icrt(ntt).download(), a)
    assert np.array_equal(ctx.icrt(ctx.upload(a)).download(), oracle.icrt(G, a))


@pytest.mark.parametrize("B,L,b,K", [(1 << 15, 5, 2, 15), (1 << 16, 4, 2, 16), (1 << 16, 4, 4, 8), (10485760000, 8, 38, 7)])
def test_decompositions(ctx, oracle, B, L, b, K):
    if B ** L < P:
        pytest.skip("uniform coefficients do not fit L digits of base B in this field")
    a = rand_elems(G, 37, 3)
    dec = ctx.gadget_decompose(ctx.upload(a, 1), B, L)
    exp = oracle.gadget_decompose(G, a, B, L)
    assert np.array_equal(dec.download(), exp)
    assert np.array_equal(ctx.gadget_recompose(dec, B, L).download(), a)
    if b ** K >= B:
        pieces = ctx.decompose_to_vec(dec, b, K)
        expk = oracle.decompose_to_vec(G, exp, b, K)
        for k in range(K):
            assert np.array_equal(pieces[k].download(), expk[k])


def test_decompose_does_not_fit(ctx, gpu):
    a = rand_elems(G, 4, 4)     # uniform coefficients need ~log2(p) bits
    with pytest.raises(gpu.LfError) as e:
        ctx.gadget_decompose(ctx.upload(a, 1), 16, 2)
    assert e.value.code == -9
    ok = ctx.gadget_decompose(ctx.upload(a, 1), 1 << 16, 4)     # the context stays usable
    assert len(ok) == 16


def test_fhat(ctx, oracle):
    a = rand_elems(G, 50, 5)
    a[40:] = 0
    exp, lens = oracle.fhat(G, a)
    got = ctx.fhat(ctx.upload(a, 1))
    for j in range(TAU):
        assert np.array_equal(got[j].download(), exp[j])     # (the oracle's truncated tail is zeros)


@pytest.mark.parametrize("kappa,n,count", [(1, 1, 1), (3, 100, 1), (4, 1500, 2), (5, 1024, 7), (2, 3000, 15)])
def test_commit(ctx, oracle, gpu, kappa, n, count):
    A = rand_elems(G, kappa * n, 6).reshape(kappa, n, D)
    sch = gpu.AjtaiCommitmentScheme(ctx, A)
    assert (sch.kappa(), sch.width()) == (kappa, n)
    fs = [rand_elems(G, n, 7 + i) for i in range(count)]
    got = sch.commit_batch([ctx.upload(f) for f in fs])
    for i in range(count):
        assert np.array_equal(got[i], oracle.commit(G, A, fs[i]))
    assert np.array_equal(sch.commit(ctx.upload(fs[0])), got[0])


def test_commit_closed_form_and_errors(ctx, gpu):
    # commitment_scheme.rs:142-160
    n, kappa = 1 << 15, 9
    A = np.zeros((kappa, n, D), dtype=np.uint64)
    A[:, :, ::TAU] = (np.arange(kappa, dtype=np.uint64)[:, None] * np.uint64(n) + np.arange(n, dtype=np.uint64)[None, :])[:, :, None]
    f = np.zeros((n, D), dtype=np.uint64); f[:, ::TAU] = 2
    sch = gpu.AjtaiCommitmentScheme(ctx, A)
    cm = sch.commit(ctx.upload(f))
    for i in range(kappa):
        exp = const_el(n * (2 * i * n + (n - 1)))
        assert np.array_equal(cm[i], exp)
    with pytest.raises(gpu.LfError) as e:       # commitment_scheme.rs:37-44
        sch.commit(ctx.upload(f[:100]))
    assert e.value.code == -1


def test_commit_linearity(ctx, gpu):
    n, kappa = 4096, 3
    A = rand_elems(G, kappa * n, 8).reshape(kappa, n, D)
    sch = gpu.AjtaiCommitmentScheme(ctx, A)
    f, g = rand_elems(G, n, 9), rand_elems(G, n, 10)
    fg = ((f.astype(object) + g.astype(object)) % P).astype(np.uint64)
    cf, cg, cfg = sch.commit(ctx.upload(f)), sch.commit(ctx.upload(g)), sch.commit(ctx.upload(fg))
    assert np.array_equal(((cf.astype(object) + cg.astype(object)) % P).astype(np.uint64), cfg)


def test_spmv(ctx, oracle, gpu):
    rng = np.random.default_rng(3)
    nrows, ncols = 70, 40
    counts = rng.integers(0, 4, nrows); counts[-5:] = 0
    row_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
    nnz = int(row_ptr[-1])
    M = dict(nrows=nrows, ncols=ncols, row_ptr=row_ptr, col=rng.integers(0, ncols, nnz).astype(np.uint64), val=rand_elems(G, nnz, 11))
    z = rand_elems(G, ncols, 12)
    sm = gpu.SparseMatrix(ctx, M)
    assert np.array_equal(sm.mat_vec_mul(ctx.upload(z)).download(), oracle.spmv(G, M, z))
    with pytest.raises(gpu.LfError) as e:
        sm.mat_vec_mul(ctx.upload(z[:10]))
    assert e.value.code == -2


@pytest.mark.parametrize("s", [1, 2, 7, 11])
def test_eq_table(ctx, oracle, s):
    r = rand_elems(G, s, 13)
    assert np.array_equal(ctx.eq_table(r).download(), oracle.eq_table(G, r))


@pytest.mark.parametrize("nv,lens", [(6, [64, 64, 64]), (7, [128, 100, 3, 0, 77]), (3, [8] * 9)])
def test_evaluate_mles(ctx, oracle, gpu, nv, lens):
    point = rand_sf_broadcast(G, nv, 14)
    full = 1 << nv
    mles = [rand_elems(G, l, 15 + i) for i, l in enumerate(lens)]
    # batches need equal lengths on the device side: pad with explicit zeros (same MLE)
    padded = np.zeros((len(lens), full, D), dtype=np.uint64)
    for i, m in enumerate(mles):
        padded[i, :len(m)] = m
    exp = oracle.evaluate_mles(G, padded, nv, point)
    got = ctx.evaluate_mles([ctx.upload(p) for p in padded], nv, point)
    assert np.array_equal(got, exp)
    for i, m in enumerate(mles):      # truncated tables: implicit zero tail
        assert np.array_equal(ctx.evaluate_mles([ctx.upload(m)], nv, point)[0], exp[i])
    with pytest.raises(gpu.LfError) as e:
        ctx.evaluate_mles([ctx.upload(padded[0])], nv, point[:-1])
    assert e.value.code == -3


def test_lincomb(ctx, oracle):
    n, cnt = 300, 35
    vecs = [rand_elems(G, n, 30 + i) for i in range(cnt)]
    coef = rand_elems(G, cnt, 99)
    acc = np.zeros((n, D), dtype=object)
    for c, v in zip(coef, vecs):
        acc = (acc + oracle.ntt_mul(G, np.ascontiguousarray(np.broadcast_to(c, (n, D))), v).astype(object)) % P
    got = ctx.lincomb(coef, [ctx.upload(v) for v in vecs]).download()
    assert np.array_equal(got, acc.astype(np.uint64))


@pytest.mark.parametrize("nv,M,deg,idx", [(5, 3, 3, [[0, 1, 2], [1, 1]]), (1, 2, 2, [[0, 1]]), (6, 8, 4, [[0, 1, 2, 3], [4, 5], [6], [7, 7, 7]])])
def test_sumcheck_products(ctx, oracle, gpu, nv, M, deg, idx):
    mles = rand_elems(G, M * (1 << nv), 40).reshape(M, 1 << nv, D)
    comb = dict(kind="products", coef=rand_elems(G, len(idx), 41), idx=idx)
    emsgs, epoint, efinal = oracle.sumcheck_prove(G, oracle.transcript(G), mles, nv, deg, comb, want_final=True)
    msgs, point, final = gpu.MLSumcheck.prove_as_subprotocol(ctx, gpu.Transcript(G), [ctx.upload(m) for m in mles], nv, deg, comb, want_final=True)
    assert np.array_equal(msgs, emsgs) and np.array_equal(point, epoint) and np.array_equal(final, efinal)


def test_sumcheck_lin_truncated(ctx, oracle, gpu):
    nv = 6
    lens = [40, 64, 17, 64]
    mles = np.zeros((4, 64, D), dtype=np.uint64)
    for i, l in enumerate(lens):
        mles[i, :l] = rand_elems(G, l, 50 + i)
    one = const_el(1)
    neg = const_el(P - 1)
    comb = dict(kind="lin", coef=np.stack([one, neg]), idx=[[0, 1], [2]])
    emsgs, epoint = oracle.sumcheck_prove(G, oracle.transcript(G), mles, nv, 3, comb, lens=lens)
    msgs, point = gpu.MLSumcheck.prove_as_subprotocol(ctx, gpu.Transcript(G), [ctx.upload(mles[i, :l]) for i, l in enumerate(lens)], nv, 3, comb)
    assert np.array_equal(msgs, emsgs) and np.array_equal(point, epoint)


@pytest.mark.parametrize("nv,K2,digits", [(4, 2, True), (5, 6, True), (3, 4, False), (1, 2, True)])
def test_sumcheck_fold(ctx, oracle, gpu, nv, K2, digits):
    """FOLD comb (folding/utils.rs:273-325) through the generic entry point: f-hat tables given as NTT vectors."""
    n = 1 << nv
    M = 5 + K2 * TAU
    mles = np.zeros((M, n, D), dtype=np.uint64)
    mles[:5] = rand_elems(G, 5 * n, 60).reshape(5, n, D)
    rng = np.random.default_rng(5)
    if digits:      # what the protocol feeds: balanced digits embedded per slot, with zero entries to hit the skips
        dg = rng.integers(-1, 2, (K2 * TAU, n, synth.RINGS[G]["S"]))
        mles[5:, :, ::TAU] = np.where(dg < 0, P - 1, dg).astype(np.uint64)
    else:
        mles[5:] = rand_elems(G, K2 * TAU * n, 61).reshape(K2 * TAU, n, D)
    mu = rand_sf_broadcast(G, K2, 62)
    comb = dict(kind="fold", mu=mu, b=2)
    emsgs, epoint, efinal = oracle.sumcheck_prove(G, oracle.transcript(G), mles, nv, 4, comb, want_final=True)
    msgs, point, final = gpu.MLSumcheck.prove_as_subprotocol(ctx, gpu.Transcript(G), [ctx.upload(m) for m in mles], nv, 4, comb, want_final=True)
    assert np.array_equal(msgs, emsgs) and np.array_equal(point, epoint) and np.array_equal(final, efinal)


def test_sumcheck_misuse(ctx, gpu):
    with pytest.raises(gpu.LfError):
        gpu.MLSumcheck.prove_as_subprotocol(ctx, gpu.Transcript(G), [ctx.upload(rand_elems(G, 1, 1))], 0, 2, dict(kind="products", coef=rand_elems(G, 1, 2), idx=[[0]]))


CASES = [  # W, B, L, b, K, kappa, kind     (decomposition_parameters.rs:49-113 and benches/config.toml rows)
    (4, 1 << 15, 5, 2, 15, 4, "scalar"),
    (4, 1 << 15, 5, 2, 15, 4, "non_scalar"),
    (8, 1 << 16, 4, 2, 16, 3, "uniform"),
    (4, 1024, 2, 2, 10, 4, "scalar"),
    (64, 1 << 16, 4, 2, 16, 5, "non_scalar"),
    (256, 1 << 13, 5, 2, 13, 6, "uniform"),
    (8, 1 << 8, 4, 2, 8, 4, "non_scalar"),      # BabyBearDP
    (4, 1 << 8, 8, 2, 10, 4, "uniform"),        # FrogDP
]


@pytest.mark.parametrize("W,B,L,b,K,kappa,kind", CASES)
def test_nifs_prove_matches_oracle(ctx, oracle, oracle_ops, gpu, W, B, L, b, K, kappa, kind):
    """One full prover step: proof, folded LCCCS and folded witness byte-identical to the oracle; the oracle's verifier
    accepts the GPU proof."""
    if kind != "scalar" and B ** L < P:
        pytest.skip("witness coefficients do not fit L digits of base B in this field")
    if G != GOLD and W > 64:
        pytest.skip("large case kept for the tuned ring only (the generic rings are parity-only and slow)")
    prob = synth.make_instance(G, W, B, L, b, K, kappa, kind=kind, config_id=2, ops=oracle_ops)
    # the GPU's own witness / commitment / accumulator construction agrees with the oracle's
    assert np.array_equal(ctx.witness_f_from_w_ccs(G, prob["w_ccs"], B, L), prob["w_i_f"])
    assert np.array_equal(ctx.commit(G, prob["A"], prob["w_i_f"]), prob["cm_i_cm"])
    pr = gpu.NIFSProver(ctx, prob)
    lc_lin, pf_lin = pr.linearize(prob, gpu.Transcript(G))
    elc_lin, epf_lin = oracle.linearize(prob, oracle.transcript(G))
    assert np.array_equal(pf_lin, epf_lin) and np.array_equal(lc_lin, elc_lin)
    eproof, elc, ef, _ = oracle.nifs_prove(prob, oracle.transcript(G))
    proof, lc, f = pr.prove(prob, gpu.Transcript(G))
    assert np.array_equal(proof, eproof)
    assert np.array_equal(lc, elc)
    assert np.array_equal(f, ef)
    oracle.nifs_verify(prob, oracle.transcript(G), proof)
    # resident-witness entry point gives the same step
    wa, wi = pr.upload_witness(prob["w_acc_f"]), pr.upload_witness(prob["w_i_f"])
    proof2, lc2, w = pr.prove_resident(prob, wa, wi, gpu.Transcript(G), keep_witness=True)
    assert np.array_equal(proof2, eproof) and np.array_equal(lc2, elc) and np.array_equal(pr.download_witness(w), ef)
    for h in (wa, wi, w):
        pr.free_witness(h)
    pr.close()


@pytest.mark.parametrize("W,B,L,b,K,kappa,kind", [(8, 1 << 8, 4, 2, 8, 4, "non_scalar"), (4, 1 << 15, 5, 2, 15, 3, "scalar"), (16, 1 << 8, 4, 2, 8, 8, "uniform")])
def test_nifs_degree_three_ccs_matches_oracle(ctx, oracle, oracle_ops, gpu, W, B, L, b, K, kappa, kind):
    """BASELINE configs[2] shape: the reference's degree-three CCS (arith/ccs.rs:14-43; t = 4 matrices, LIN sumcheck of degree 4)
    through the full step, byte-identical to the oracle, on every ring (BabyBearDP = (256, 4, 2, 8), kappa = 8 for the last case)."""
    if kind != "scalar" and B ** L < P:
        pytest.skip("witness coefficients do not fit L digits of base B in this field")
    prob = synth.make_instance(G, W, B, L, b, K, kappa, kind=kind, config_id=5, ops=oracle_ops, degree=3)
    assert prob["ccs"]["t"] == 4 and prob["ccs"]["d"] == 3
    pr = gpu.NIFSProver(ctx, prob)
    eproof, elc, ef, _ = oracle.nifs_prove(prob, oracle.transcript(G))
    proof, lc, f = pr.prove(prob, gpu.Transcript(G))
    assert np.array_equal(proof, eproof) and np.array_equal(lc, elc) and np.array_equal(f, ef)
    oracle.nifs_verify(prob, oracle.transcript(G), proof)
    pr.close()


def test_ops_ntt_mul_matches_oracle(ctx, oracle):
    a, b = rand_elems(G, 9, 41), rand_elems(G, 9, 42)
    assert np.array_equal(ctx.ntt_mul(G, a, b), oracle.ntt_mul(G, a, b))


def _lfplus_setchk_terms(n_M, ncols, n_m):
    """term list of the LatticeFold+ set-check batch (latticefold-plus/src/setchk.rs:155-186):
    sum_i rc^i eq_i sum_j alpha_i^j (m_ij^2 - m'_ij)  +  sum_i rc^(n_M+i) alpha eq (m^2 - m'); tables laid out as the reference pushes them"""
    terms = []
    for i in range(n_M):
        s = i * (2 * ncols + 1)
        for j in range(ncols):
            terms += [[s + 2 * ncols, s + 2 * j, s + 2 * j], [s + 2 * ncols, s + 2 * j + 1]]
    base = n_M * (2 * ncols + 1)
    for i in range(n_m):
        s = base + 3 * i
        terms += [[s + 2, s, s], [s + 2, s + 1]]
    return terms, base + 3 * n_m


@pytest.mark.parametrize("nv,shape", [(5, "r1cs"), (4, "setchk"), (6, "cm"), (3, "wide")])
def test_sumcheck_general_terms_lfplus_shapes(ctx, oracle, gpu, nv, shape):
    """sums of products beyond the register-resident kernel's 8 tables / 4 terms (k_sc_terms): the comb functions of LatticeFold+
    -- v0 (v1 v2 - v3) (r1cs.rs:92), the set-check batch (setchk.rs:155-186), the commitment-transformation batch (cm.rs:285-307) --
    and a wide random sum of products, against the oracle's sumcheck"""
    if shape == "r1cs":
        M, deg, idx = 4, 3, [[0, 1, 2], [0, 3]]
        coef = np.stack([const_el(1), const_el(P - 1)])
    elif shape == "setchk":
        idx, M = _lfplus_setchk_terms(2, 3, 2); deg = 3
        coef = rand_sf_broadcast(G, len(idx), 71)
    elif shape == "cm":
        L_, Mlen = 2, 2; M = 1 + L_ * (4 + 4 * Mlen) + 2; deg = 2; idx = []
        for l in range(L_):
            li = 1 + l * (4 + 4 * Mlen)
            idx += [[0, li + q] for q in range(4 + 4 * Mlen)] + [[li, M - 2], [li, M - 1]]
        coef = rand_sf_broadcast(G, len(idx), 72)
    else:
        M, deg = 19, 4
        rng = np.random.default_rng(9)
        idx = [list(rng.integers(0, M, int(rng.integers(1, 5)))) for _ in range(23)]
        coef = rand_elems(G, len(idx), 73)
    mles = rand_elems(G, M * (1 << nv), 70).reshape(M, 1 << nv, D)
    comb = dict(kind="products", coef=coef, idx=idx)
    emsgs, epoint, efinal = oracle.sumcheck_prove(G, oracle.transcript(G), mles, nv, deg, comb, want_final=True)
    msgs, point, final = gpu.MLSumcheck.prove_as_subprotocol(ctx, gpu.Transcript(G), [ctx.upload(m) for m in mles], nv, deg, comb, want_final=True)
    assert np.array_equal(msgs, emsgs) and np.array_equal(point, epoint) and np.array_equal(final, efinal)


@pytest.mark.parametrize("ring", [synth.RING_GOLDILOCKS, synth.RING_BABYBEAR])
def test_montgomery_representation_at_the_boundary(gpu, oracle, ring):
    """LF_REPR_MONTGOMERY: witness-sized host vectors given as ark-ff Montgomery limbs (a * 2^64 mod p) are the same device vectors as
    their canonical images, in both directions, and a commitment of Montgomery-form inputs equals the canonical one"""
    R = synth.RINGS[ring]; p, d = R["p"], R["d"]
    c = gpu.Context(ring, 0)
    try:
        a = rand_elems(ring, 100, 91)
        mont = ((a.astype(object) << 64) % p).astype(np.uint64)
        c.set_bulk_repr(True); v = c.upload(mont); c.set_bulk_repr(False)
        assert np.array_equal(v.download(), a)
        c.set_bulk_repr(True); got = v.download(); c.set_bulk_repr(False)
        assert np.array_equal(got, mont)
        kappa, n = 3, 100
        A = rand_elems(ring, kappa * n, 92).reshape(kappa, n, d)
        want = oracle.commit(ring, A, a)
        c.set_bulk_repr(True)
        sch = gpu.AjtaiCommitmentScheme(c, ((A.astype(object) << 64) % p).astype(np.uint64))
        cm = sch.commit(c.upload(mont))                    # the kappa result elements are small data: canonical in either mode
        c.set_bulk_repr(False)
        assert np.array_equal(cm, want)
        del sch
    finally:
        c.close()
