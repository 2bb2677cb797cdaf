#!/usr/bin/env python3
"""Golden digests of whole prover steps: tests/golden/step_digests.json.

For each case of tests/helpers.py: STEP_GOLDEN_CASES (ring, W, B, L, b, K, kappa, kind, CCS degree, config_id) the deterministic synthetic instance of
latticefold_b200/synth.py is proved by the CPU oracle (oracle/, NIFSProver::prove, nifs.rs:48-103) and the SHA-256 of the proof, of
the folded LCCCS and of the folded witness (little-endian u64 limbs) is recorded.  The fixture pins the oracle and the instance
generator against silent changes (tests/test_oracle_protocol.py) and gives the CUDA path a committed answer to hit on the GPU box
without trusting the oracle built there (tests/test_gpu_parity.py).  The oracle itself is pinned by the reference's KATs
(tests/test_oracle_kats.py); the reference cannot be run (no Rust toolchain), so these are oracle outputs, not reference outputs.
Needs no GPU and no reference tree:  python tools/make_step_golden.py
"""
import json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Oracle
from tests.helpers import OracleOps, STEP_GOLDEN_CASES, STEP_GOLDEN_PATH, step_case_key, step_digests, step_instance

if __name__ == "__main__":
    oracle = Oracle(); ops = OracleOps(oracle)
    out = {"source": "oracle/ (CPU restatement of nifs.rs:48-103) on latticefold_b200/synth.py instances; tools/make_step_golden.py",
           "hash": "sha256 over little-endian u64 limbs", "cases": {}}
    for c in STEP_GOLDEN_CASES:
        prob = step_instance(c, ops)
        proof, lc, f, _ = oracle.nifs_prove(prob, oracle.transcript(c[0]))
        out["cases"][step_case_key(*c)] = step_digests(proof, lc, f)
    json.dump(out, open(STEP_GOLDEN_PATH, "w"), indent=1)
    print("wrote", STEP_GOLDEN_PATH, len(out["cases"]), "cases")
