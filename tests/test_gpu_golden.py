"""The CUDA path (through the C ABI) against committed golden fixtures: SHA-256 digests of whole prover steps produced by the CPU
oracle on the deterministic synthetic instances (tests/golden/step_digests.json, tools/make_step_golden.py).  Own module so that no
other context is alive on the device while these run."""
import pytest

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
