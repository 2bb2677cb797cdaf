"""Protocol-level acceptance of the CPU oracle: sumcheck prove->verify, the full
NIFS prove->verify round trip at the reference's toy sizes, and tamper rejection.  CPU only.
Mirrors crates/latticefold/src/utils/sumcheck.rs:148-224 and crates/latticefold/src/nifs/tests.rs:58-117."""
import numpy as np
import pytest

from latticefold_b200 import synth
from oracle.pyoracle import OracleError
from tests import helpers
from tests.helpers import rand_elems

G, BB, FROG = synth.RING_GOLDILOCKS, synth.RING_BABYBEAR, synth.RING_FROG

CASES = [  # ring, W, B, L, b, K, kappa, kind        (DPs: decomposition_parameters.rs:49-113)
    (G, 4, 1 << 15, 5, 2, 15, 4, "scalar"),
    (G, 4, 1 << 15, 5, 2, 15, 4, "non_scalar"),
    (G, 8, 1 << 16, 4, 2, 16, 3, "uniform"),
    (G, 4, 1024, 2, 2, 10, 4, "scalar"),
    (BB, 4, 1 << 8, 4, 2, 8, 4, "non_scalar"),
    (FROG, 4, 1 << 8, 8, 2, 10, 4, "uniform"),
]


@pytest.mark.parametrize("ring", [G, BB, FROG])
def test_sumcheck_roundtrip(oracle, ring):
    R = synth.RINGS[ring]; d = R["d"]; nv, M = 4, 3
    mles = rand_elems(ring, M * (1 << nv), 7).reshape(M, 1 << nv, d)
    comb = dict(kind="products", coef=rand_elems(ring, 2, 8), idx=[[0, 1, 2], [1, 1]])
    msgs, point = oracle.sumcheck_prove(ring, oracle.transcript(ring), mles, nv, 3, comb)
    p = R["p"]
    claimed = ((msgs[0, 0].astype(object) + msgs[0, 1].astype(object)) % p).astype(np.uint64)
    exp, vpoint = oracle.sumcheck_verify(ring, oracle.transcript(ring), nv, 3, claimed, msgs)
    assert np.array_equal(point, vpoint)
    # verifier's final claim == comb(mle_*(r))
    pt_ring = np.ascontiguousarray(np.broadcast_to(point[:, None, :], (nv, R["S"], R["tau"])).reshape(nv, d))
    vals = oracle.evaluate_mles(ring, mles, nv, pt_ring)
    t0 = oracle.ntt_mul(ring, oracle.ntt_mul(ring, oracle.ntt_mul(ring, comb["coef"][0:1], vals[0:1]), vals[1:2]), vals[2:3])[0]
    t1 = oracle.ntt_mul(ring, oracle.ntt_mul(ring, comb["coef"][1:2], vals[1:2]), vals[1:2])[0]
    assert np.array_equal(((t0.astype(object) + t1.astype(object)) % p).astype(np.uint64), exp)
    # wrong sum rejects (sumcheck.rs:197-224)
    bad = claimed.copy(); bad[0] = (int(bad[0]) + 1) % p
    with pytest.raises(OracleError):
        oracle.sumcheck_verify(ring, oracle.transcript(ring), nv, 3, bad, msgs)


@pytest.mark.parametrize("ring,W,B,L,b,K,kappa,kind", CASES)
def test_nifs_prove_verify(oracle, oracle_ops, ring, W, B, L, b, K, kappa, kind):
    prob = synth.make_instance(ring, W, B, L, b, K, kappa, kind=kind, config_id=1, ops=oracle_ops)
    proof, lc, f, _ = oracle.nifs_prove(prob, oracle.transcript(ring))
    lc_v = oracle.nifs_verify(prob, oracle.transcript(ring), proof)
    assert np.array_equal(lc, lc_v)
    # the folded commitment opens to the folded witness: cm_0 == A * f_0  (homomorphism, folding/utils.rs:460-521)
    out = synth.split_lcccs(ring, prob, lc)
    assert np.array_equal(out["cm"], oracle.commit(ring, prob["A"], f))
    # tamper: flip one limb somewhere in each sub-proof region -> reject (linearization/tests/mod.rs:364-396 etc.)
    for pos in (0, proof.size // 3, proof.size // 2, proof.size - 1):
        bad = proof.copy(); bad[pos] = (int(bad[pos]) + 1) % synth.RINGS[ring]["p"]
        with pytest.raises(OracleError):
            oracle.nifs_verify(prob, oracle.transcript(ring), bad)


def test_witness_forms_agree(oracle):
    # arith.rs:516-548: from_w_ccs / from_f / from_f_coeff describe the same witness
    B, L = 1 << 15, 5
    w = rand_elems(G, 6, 3)
    fc = oracle.gadget_decompose(G, oracle.icrt(G, w), B, L)
    f = oracle.crt(G, fc)
    assert np.array_equal(oracle.icrt(G, f), fc)
    assert np.array_equal(oracle.gadget_recompose(G, f, B, L), w)
    assert np.array_equal(oracle.crt(G, oracle.gadget_recompose(G, fc, B, L)), w)


@pytest.mark.parametrize("ring,W,B,L,b,K,kappa,kind", [(BB, 4, 1 << 8, 4, 2, 8, 4, "non_scalar"), (G, 4, 1 << 15, 5, 2, 15, 3, "uniform"), (FROG, 4, 1 << 8, 8, 2, 10, 3, "scalar")])
def test_nifs_prove_verify_degree_three_ccs(oracle, oracle_ops, ring, W, B, L, b, K, kappa, kind):
    """the reference's dummy degree-three CCS (arith/ccs.rs:14-43) through prove -> verify (BASELINE configs[2] shape)"""
    prob = synth.make_instance(ring, W, B, L, b, K, kappa, kind=kind, config_id=6, ops=oracle_ops, degree=3)
    assert prob["ccs"]["t"] == 4 and prob["ccs"]["q"] == 2 and prob["ccs"]["d"] == 3
    proof, lc, f, _ = oracle.nifs_prove(prob, oracle.transcript(ring))
    assert np.array_equal(lc, oracle.nifs_verify(prob, oracle.transcript(ring), proof))
    bad = proof.copy(); bad[3] = (int(bad[3]) + 1) % synth.RINGS[ring]["p"]
    with pytest.raises(OracleError):
        oracle.nifs_verify(prob, oracle.transcript(ring), bad)


@pytest.mark.parametrize("case", helpers.STEP_GOLDEN_CASES, ids=lambda c: helpers.step_case_key(*c))
def test_step_digests_match_committed_golden(oracle, oracle_ops, case):
    """the oracle's whole prover step (proof, folded LCCCS, folded witness) on the deterministic synthetic instance still hashes to
    tests/golden/step_digests.json: pins oracle + instance generator; the GPU path is held to the same file in test_gpu_parity.py"""
    prob = helpers.step_instance(case, oracle_ops)
    proof, lc, f, _ = oracle.nifs_prove(prob, oracle.transcript(case[0]))
    assert helpers.step_digests(proof, lc, f) == helpers.step_golden()[helpers.step_case_key(*case)]


@pytest.mark.parametrize("config,log_w", [("c2", 10), ("c3", 8)])
def test_bench_digests_match_committed_golden(oracle, config, log_w):
    """the oracle on the instance bench.py proves (synth.bench_instance; small sizes here, the full sizes are regenerated by
    tools/make_bench_golden.py) still hashes to tests/golden/bench_digests.json"""
    from tools.make_bench_golden import oracle_bench_problem
    wl, prob = oracle_bench_problem(oracle, config, log_w)
    proof, lc, f, _ = oracle.nifs_prove(prob, oracle.transcript(wl["ring"]))
    gold = helpers.bench_golden()[helpers.bench_case_key(config, log_w)]
    got = helpers.step_digests(proof, lc, f)
    assert all(got[k] == gold[k] for k in ("proof", "lcccs", "witness", "proof_words"))


def test_c3lin_digest_matches_committed_golden(oracle):
    """commit + linearization of the BASELINE configs[2] workload (small size) still hashes to tests/golden/bench_digests.json, and
    the oracle's and the product's LFLinearizationVerifier accept it"""
    import latticefold_b200 as lf
    from tools.make_bench_golden import oracle_c3lin, c3lin_digests
    wl, prob, lc, pf, _ = oracle_c3lin(oracle, 10)
    gold = helpers.bench_golden()[helpers.bench_case_key("c3lin", 10)]
    got = c3lin_digests(prob["cm_i_cm"], lc, pf)
    assert all(got[k] == gold[k] for k in got)
    assert np.array_equal(oracle.linearization_verify(prob, oracle.transcript(wl["ring"]), pf), lc)
    assert np.array_equal(lf.linearization_verify(prob, lf.Transcript(wl["ring"]), pf), lc)
    bad = pf.copy(); bad[3] = (int(bad[3]) + 1) % synth.RINGS[wl["ring"]]["p"]
    with pytest.raises(OracleError):
        oracle.linearization_verify(prob, oracle.transcript(wl["ring"]), bad)
    with pytest.raises(lf.LfError):
        lf.linearization_verify(prob, lf.Transcript(wl["ring"]), bad)
