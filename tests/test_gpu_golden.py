"""The CUDA path (through the C ABI) against committed golden fixtures: SHA-256 digests of whole prover steps produced by the CPU
oracle on the deterministic synthetic instances (tests/golden/step_digests.json, tools/make_step_golden.py).  Own module so that no
other context is alive on the device while these run."""
import numpy as np
import pytest

from latticefold_b200 import synth
from tests import helpers

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", helpers.STEP_GOLDEN_CASES, ids=lambda c: helpers.step_case_key(*c))
def test_nifs_prove_matches_committed_golden(oracle_ops, gpu, case):
    """the CUDA path against the committed digests of whole prover steps (tests/golden/step_digests.json, tools/make_step_golden.py):
    proof, folded LCCCS and folded witness, through both entry points (host buffers and resident witnesses)"""
    ring = case[0]
    want = helpers.step_golden()[helpers.step_case_key(*case)]
    prob = helpers.step_instance(case, oracle_ops)
    c = gpu.Context(ring, 0)
    try:
        pr = gpu.NIFSProver(c, prob)
        proof, lc, f = pr.prove(prob, gpu.Transcript(ring))
        assert helpers.step_digests(proof, lc, f) == want
        if case[8] == 2:   # the resident-witness entry point on the R1CS cases
            wa, wi = pr.upload_witness(prob["w_acc_f"]), pr.upload_witness(prob["w_i_f"])
            proof2, lc2, w = pr.prove_resident(prob, wa, wi, gpu.Transcript(ring), keep_witness=True)
            assert helpers.step_digests(proof2, lc2, pr.download_witness(w)) == want
            for h in (wa, wi, w):
                pr.free_witness(h)
        pr.close()
    finally:
        c.close()


def _gpu_bench_problem(gpu, ctx, config, log_w):
    """the instance bench.py proves on one GPU, completed by the product library exactly as bench.py completes it"""
    wl = synth.bench_workload(config, log_w)
    prob = synth.bench_instance(wl, 0, 1, ops=ctx)
    f = ctx.witness_f_from_w_ccs(wl["ring"], prob["w_ccs"], wl["B"], wl["L"])
    prob["w_i_f"] = prob["w_acc_f"] = f
    prob["cm_i_cm"] = np.ascontiguousarray(ctx.commit(wl["ring"], prob["A"], f))
    prob["acc"] = ctx.linearize(prob)
    return wl, prob


@pytest.mark.parametrize("config,log_w", [("c2", 10), ("c2", 12), ("c2", 16), ("c3", 8), ("c3", 10), ("c3", 12), ("c3", 16)])
def test_bench_step_matches_committed_golden(gpu, config, log_w):
    """the step bench.py times (BASELINE configs[1] at log_w = 16; configs[2]'s shape on the BabyBear ring) against the digests the
    CPU oracle produced for the same instance (tests/golden/bench_digests.json, tools/make_bench_golden.py): proof, folded LCCCS and
    folded witness byte for byte -- including the instance set-up (witness, commitment and accumulator come from the GPU here and
    from the oracle there)"""
    key = helpers.bench_case_key(config, log_w)
    gold = helpers.bench_golden()
    if key not in gold:
        pytest.skip("no committed digest for " + key)
    ring = synth.bench_workload(config, log_w)["ring"]
    c = gpu.Context(ring, 0)
    try:
        wl, prob = _gpu_bench_problem(gpu, c, config, log_w)
        pr = gpu.NIFSProver(c, prob)
        proof, lc, f = pr.prove(prob, gpu.Transcript(ring))
        got = helpers.step_digests(proof, lc, f)
        assert {k: got[k] for k in ("proof", "lcccs", "witness", "proof_words")} == {k: gold[key][k] for k in ("proof", "lcccs", "witness", "proof_words")}
        assert np.array_equal(gpu.nifs_verify(prob, gpu.Transcript(ring), proof), lc)
        pr.close()
    finally:
        c.close()


@pytest.mark.parametrize("log_w", [10, 12, 16, 18])
def test_c3_commit_and_linearization_match_committed_golden(gpu, log_w):
    """BASELINE configs[2] as bench.py --config c3 runs it (BabyBear ring, packed 4-byte limb planes, degree-three CCS): the witness
    commitment A f and LFLinearizationProver::prove against the digests of the CPU oracle on the same instance, through the
    resident-witness entry points, plus the product's LFLinearizationVerifier"""
    key = helpers.bench_case_key("c3lin", log_w)
    gold = helpers.bench_golden()
    if key not in gold:
        pytest.skip("no committed digest for " + key)
    wl = synth.bench_workload("c3", log_w); ring = wl["ring"]
    c = gpu.Context(ring, 0)
    try:
        prob = synth.bench_instance(wl, 0, 1, ops=c)
        pr = gpu.NIFSProver(c, prob)
        f = c.witness_f_from_w_ccs(ring, prob["w_ccs"], wl["B"], wl["L"])
        w = pr.upload_witness(f)
        cm = pr.witness_commit(w); prob["cm_i_cm"] = np.ascontiguousarray(cm)
        lc, pf = pr.linearize_resident(prob, w, gpu.Transcript(ring))
        got = {"cm": helpers.limb_digest(cm), "lcccs": helpers.limb_digest(lc), "lin_proof": helpers.limb_digest(pf), "lin_proof_words": int(pf.size)}
        assert got == {k: gold[key][k] for k in got}
        assert np.array_equal(gpu.linearization_verify(prob, gpu.Transcript(ring), pf), lc)
        bad = pf.copy(); bad[5] = (int(bad[5]) + 1) % synth.RINGS[ring]["p"]
        with pytest.raises(gpu.LfError):
            gpu.linearization_verify(prob, gpu.Transcript(ring), bad)
        pr.free_witness(w); pr.close()
    finally:
        c.close()
