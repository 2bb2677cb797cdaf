"""Shared test helpers (CPU oracle adapters, random ring data)."""
import hashlib
import json
import os

import numpy as np

from latticefold_b200 import synth


class OracleOps:
    """`ops` adapter for synth.make_instance backed by the CPU oracle (used where no GPU is present)."""

    def __init__(self, oracle):
        self.o = oracle

    def witness_f_from_w_ccs(self, ring, w_ccs, B, L):  # Witness::from_w_ccs, arith.rs:230-248
        return self.o.crt(ring, self.o.gadget_decompose(ring, self.o.icrt(ring, w_ccs), B, L))

    def commit(self, ring, A, f):
        return self.o.commit(ring, A, f)

    def ntt_mul(self, ring, a, b):
        return self.o.ntt_mul(ring, a, b)

    def linearize(self, prob):
        lc, _ = self.o.linearize(prob, self.o.transcript(prob["ring"]))
        return synth.split_lcccs(prob["ring"], prob, lc)


def rand_elems(ring, count, seed):
    R = synth.RINGS[ring]
    return synth.uniform_field(R["p"], count * R["d"], seed).reshape(count, R["d"])


def rand_sf_broadcast(ring, count, seed):
    """ring elements that are the broadcast of one slot-field element (what sumcheck challenges look like)."""
    R = synth.RINGS[ring]
    sf = synth.uniform_field(R["p"], count * R["tau"], seed).reshape(count, 1, R["tau"])
    return np.ascontiguousarray(np.broadcast_to(sf, (count, R["S"], R["tau"])).reshape(count, R["d"]))


# ---- committed digests of whole prover steps (tests/golden/step_digests.json, written by tools/make_step_golden.py)
STEP_GOLDEN_CASES = [  # ring, W, B, L, b, K, kappa, kind, CCS degree, config_id
    (synth.RING_GOLDILOCKS, 4, 1 << 15, 5, 2, 15, 4, "scalar", 2, 2),
    (synth.RING_GOLDILOCKS, 4, 1 << 15, 5, 2, 15, 4, "non_scalar", 2, 2),
    (synth.RING_GOLDILOCKS, 8, 1 << 16, 4, 2, 16, 3, "uniform", 2, 2),
    (synth.RING_GOLDILOCKS, 64, 1 << 16, 4, 2, 16, 5, "non_scalar", 2, 2),
    (synth.RING_GOLDILOCKS, 4, 1 << 15, 5, 2, 15, 3, "scalar", 3, 5),
    (synth.RING_BABYBEAR, 8, 1 << 8, 4, 2, 8, 4, "non_scalar", 2, 2),
    (synth.RING_BABYBEAR, 16, 1 << 8, 4, 2, 8, 8, "uniform", 3, 5),
    (synth.RING_FROG, 4, 1 << 8, 8, 2, 10, 4, "uniform", 2, 2),
]
STEP_GOLDEN_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "step_digests.json")


def step_case_key(ring, W, B, L, b, K, kappa, kind, degree, config_id):
    return "ring%d_W%d_B%d_L%d_b%d_K%d_k%d_%s_deg%d_cfg%d" % (ring, W, B, L, b, K, kappa, kind, degree, config_id)


def limb_digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype="<u8").tobytes()).hexdigest()


def step_digests(proof, lc, f):
    return {"proof": limb_digest(proof), "lcccs": limb_digest(lc), "witness": limb_digest(f), "proof_words": int(proof.size)}


def step_golden():
    return json.load(open(STEP_GOLDEN_PATH))["cases"]


def step_instance(case, ops):
    ring, W, B, L, b, K, kappa, kind, degree, config_id = case
    return synth.make_instance(ring, W, B, L, b, K, kappa, kind=kind, config_id=config_id, ops=ops, degree=degree)


# ---- committed digests of the steps bench.py times (tests/golden/bench_digests.json, written by tools/make_bench_golden.py)
BENCH_GOLDEN_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bench_digests.json")


def bench_case_key(config, log_w):
    return "%s_logw%d" % (config, log_w)


def bench_golden():
    return json.load(open(BENCH_GOLDEN_PATH))["cases"]
