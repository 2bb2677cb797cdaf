"""Full-size (BASELINE configs[1], C2: W = 2^16, kappa = 26, n = 2^18) checks of the CUDA path through size-independent
properties, since the CPU oracle's prover needs minutes at this size:
  * the oracle's VERIFIER (cheap: no witness-sized data) accepts the GPU proof and reproduces the folded instance,
  * the folded commitment opens to the folded witness (cm_0 == A f_0, the Ajtai homomorphism),
  * decompose -> recompose and CRT -> ICRT round trips, commit linearity, sumcheck message consistency p(0)+p(1) chain."""
import numpy as np
import pytest

from latticefold_b200 import synth

pytestmark = pytest.mark.gpu
G = synth.RING_GOLDILOCKS
P = synth.RINGS[G]["p"]
W, B, L, b, K, KAPPA = 1 << 16, 1 << 16, 4, 2, 16, 26


@pytest.fixture(scope="module")
def setup(gpu):
    ctx = gpu.Context(G, 0)
    prob = synth.make_instance(G, W, B, L, b, K, KAPPA, kind="non_scalar", config_id=11, ops=ctx)
    yield ctx, prob
    ctx.close()


def test_full_step_verifies_and_opens(setup, oracle, gpu):
    ctx, prob = setup
    pr = gpu.NIFSProver(ctx, prob)
    proof, lc, f0 = pr.prove(prob, gpu.Transcript(G))
    light = {k: v for k, v in prob.items() if k not in ("A", "w_i_f", "w_acc_f")}      # the verifier needs no witness-sized input
    lc_v = oracle.nifs_verify(light, oracle.transcript(G), proof)
    assert np.array_equal(lc_v, lc)
    assert np.array_equal(gpu.nifs_verify(prob, gpu.Transcript(G), proof), lc)      # the product's own (host) verifier agrees
    out = synth.split_lcccs(G, prob, lc)
    sch = gpu.AjtaiCommitmentScheme(ctx, prob["A"])
    assert np.array_equal(sch.commit(ctx.upload(f0)), out["cm"])
    # a second, chained step: the folded pair becomes the accumulator (IVC hand-off) and still verifies
    prob2 = dict(prob); prob2["acc"] = out; prob2["w_acc_f"] = f0
    proof2, lc2, f1 = pr.prove(prob2, gpu.Transcript(G))
    light2 = {k: v for k, v in prob2.items() if k not in ("A", "w_i_f", "w_acc_f")}
    assert np.array_equal(oracle.nifs_verify(light2, oracle.transcript(G), proof2), lc2)
    assert np.array_equal(sch.commit(ctx.upload(f1)), synth.split_lcccs(G, prob, lc2)["cm"])
    # tampering with the GPU proof is caught
    from oracle.pyoracle import OracleError
    bad = proof.copy(); bad[bad.size // 2] ^= np.uint64(1)
    with pytest.raises(OracleError):
        oracle.nifs_verify(light, oracle.transcript(G), bad)
    pr.close()


def test_roundtrips_and_linearity_fullsize(setup, gpu):
    ctx, prob = setup
    f = prob["w_i_f"]
    n = f.shape[0]
    dv = ctx.upload(f)
    coeff = ctx.icrt(dv)
    assert np.array_equal(ctx.crt(coeff).download(), f)
    pieces = ctx.decompose_to_vec(coeff, b, K)            # witness of from_w_ccs is B-bounded, so K digits of base b fit
    fc = coeff.download()
    for k in range(K):                                    # balanced base-2 digits: -1, 0, 1
        pk = pieces[k].download()
        assert np.isin(pk, np.array([0, 1, P - 1], dtype=np.uint64)).all()
    # exact check on a sample of rows with Python integers
    rows = np.arange(0, n, max(1, n // 97))
    tot = np.zeros((len(rows), 24), dtype=object)
    for k in reversed(range(K)):
        tot = (tot * b + pieces[k].download()[rows].astype(object)) % P
    assert np.array_equal(tot.astype(np.uint64), fc[rows])
    # gadget_recompose(f, B, L) returns the CCS witness
    assert np.array_equal(ctx.gadget_recompose(dv, B, L).download(), prob["w_ccs"])
    # commit linearity at full size
    sch = gpu.AjtaiCommitmentScheme(ctx, prob["A"])
    g = np.ascontiguousarray(np.roll(f, 7, axis=0))
    fg = ((f.astype(object) + g.astype(object)) % P).astype(np.uint64)
    c1, c2, c3 = sch.commit(dv), sch.commit(ctx.upload(g)), sch.commit(ctx.upload(fg))
    assert np.array_equal(((c1.astype(object) + c2.astype(object)) % P).astype(np.uint64), c3)
    both = sch.commit_batch([dv, ctx.upload(g)])
    assert np.array_equal(both[0], c1) and np.array_equal(both[1], c2)


def test_babybear_degree_three_ccs_c3_shape(oracle, gpu):
    """BASELINE configs[2] (C3) at a size one GPU holds with 8-byte limbs: BabyBear ring, BabyBearDP (256, 4, 2, 8), kappa = 8,
    degree-three CCS, W = 2^12 (the full 2^20 needs the packed 4-byte limb layout: the folding tables alone are 174 GB at 8 bytes
    per limb).  The oracle's verifier accepts the GPU proof and the folded commitment opens to the folded witness."""
    BBR = synth.RING_BABYBEAR
    ctx = gpu.Context(BBR, 0)
    prob = synth.make_instance(BBR, 1 << 12, 256, 4, 2, 8, 8, kind="non_scalar", config_id=12, ops=ctx, degree=3)
    pr = gpu.NIFSProver(ctx, prob)
    proof, lc, f0 = pr.prove(prob, gpu.Transcript(BBR))
    light = {k: v for k, v in prob.items() if k not in ("A", "w_i_f", "w_acc_f")}
    assert np.array_equal(oracle.nifs_verify(light, oracle.transcript(BBR), proof), lc)
    sch = gpu.AjtaiCommitmentScheme(ctx, prob["A"])
    assert np.array_equal(sch.commit(ctx.upload(f0)), synth.split_lcccs(BBR, prob, lc)["cm"])
    pr.close(); ctx.close()


def test_host_buffer_step_with_pinned_memory(setup, gpu):
    """Pinned inputs make the witness uploads truly asynchronous: the accumulator's decomposition starts (auxiliary stream) while
    the incoming witness is still being copied on the main stream.  Same proof as with pageable (synchronously copied) inputs."""
    import torch
    ctx, prob = setup
    keep, p2 = [], dict(prob)
    for k in ("w_i_f", "w_acc_f"):
        t = torch.from_numpy(np.ascontiguousarray(prob[k]).view(np.int64)).pin_memory(); keep.append(t)
        p2[k] = t.numpy().view(np.uint64)
    pr = gpu.NIFSProver(ctx, prob)
    ref_proof, ref_lc, ref_f = pr.prove(prob, gpu.Transcript(G))
    for _ in range(3):
        proof, lc, f = pr.prove(p2, gpu.Transcript(G))
        assert np.array_equal(proof, ref_proof) and np.array_equal(lc, ref_lc) and np.array_equal(f, ref_f)
    pr.close()
