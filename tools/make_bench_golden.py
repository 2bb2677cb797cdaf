#!/usr/bin/env python3
"""Golden digests of the prover step bench.py times: tests/golden/bench_digests.json.

The instance is exactly what `bench.py` proves on one GPU (latticefold_b200/synth.py: bench_workload / bench_instance; BASELINE.json
configs[1] = SURVEY 8 row C2 at log_w = 16), completed the way bench.py completes it (f = Witness::from_w_ccs, cm_i = A f, accumulator =
linearization of the same instance) but with the CPU oracle doing every ring operation.  The oracle then proves the step
(NIFSProver::prove, nifs.rs:48-103) and the SHA-256 of proof, folded LCCCS and folded witness is recorded.  bench.py compares the digest
of the proof it produced on the GPU with this file after the timed region; tests/test_gpu_fullsize.py does the same under pytest.
Oracle outputs, not reference outputs (the reference cannot be built in this image).  Needs no GPU:
    python tools/make_bench_golden.py [log_w ...]        (log_w = 16 takes a few CPU-minutes and ~20 GB of RAM)
"""
import json, os, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from latticefold_b200 import synth
from oracle.pyoracle import Oracle
from tests.helpers import OracleOps, step_digests, BENCH_GOLDEN_PATH, bench_case_key


def oracle_bench_problem(orc, config, log_w):
    ops = OracleOps(orc)
    wl = synth.bench_workload(config, log_w)
    prob = synth.bench_instance(wl, 0, 1, ops=ops)
    f = np.ascontiguousarray(ops.witness_f_from_w_ccs(wl["ring"], prob["w_ccs"], wl["B"], wl["L"]))
    prob["w_i_f"] = prob["w_acc_f"] = f
    prob["cm_i_cm"] = np.ascontiguousarray(orc.commit(wl["ring"], prob["A"], f))
    prob["acc"] = ops.linearize(prob)
    return wl, prob


def oracle_c3lin(orc, log_w):
    """BASELINE configs[2] as BASELINE.md states it (C3: commit + linearization sumcheck): cm = A f, then LFLinearizationProver::prove"""
    ops = OracleOps(orc)
    wl = synth.bench_workload("c3", log_w)
    prob = synth.bench_instance(wl, 0, 1, ops=ops)
    f = np.ascontiguousarray(ops.witness_f_from_w_ccs(wl["ring"], prob["w_ccs"], wl["B"], wl["L"]))
    prob["w_i_f"] = f
    import time as _t
    t0 = _t.time()
    prob["cm_i_cm"] = np.ascontiguousarray(orc.commit(wl["ring"], prob["A"], f))
    prob.pop("A")                                   # the linearization does not read the matrix: do not let the oracle copy 19 GB of it
    lc, pf = orc.linearize(prob, orc.transcript(wl["ring"]))
    ms = (_t.time() - t0) * 1e3
    return wl, prob, lc, pf, ms


def c3lin_digests(cm, lc, pf):
    from tests.helpers import limb_digest
    return {"cm": limb_digest(cm), "lcccs": limb_digest(lc), "lin_proof": limb_digest(pf), "lin_proof_words": int(pf.size)}


if __name__ == "__main__":
    cases = [a.split(":") for a in sys.argv[1:]] or [["c2", "10"], ["c2", "12"], ["c2", "16"], ["c3", "10"], ["c3", "12"]]
    orc = Oracle()
    out = json.load(open(BENCH_GOLDEN_PATH)) if os.path.exists(BENCH_GOLDEN_PATH) else {
        "source": "oracle/ (CPU restatement of nifs.rs:48-103) on latticefold_b200/synth.py bench_instance; tools/make_bench_golden.py",
        "hash": "sha256 over little-endian u64 limbs", "cases": {}}
    for config, log_w in cases:
        log_w = int(log_w); t0 = time.time()
        if config == "c3lin":
            wl, prob, lc, pf, ms = oracle_c3lin(orc, log_w)
            orc.linearization_verify(prob, orc.transcript(wl["ring"]), pf)
            out["cases"][bench_case_key(config, log_w)] = dict(c3lin_digests(prob["cm_i_cm"], lc, pf), oracle_step_ms=round(ms, 1), oracle_threads=orc.threads())
            print(bench_case_key(config, log_w), "commit + linearization %.0f ms, total %.0f s" % (ms, time.time() - t0), flush=True)
            json.dump(out, open(BENCH_GOLDEN_PATH, "w"), indent=1)
            continue
        wl, prob = oracle_bench_problem(orc, config, log_w)
        proof, lc, f, ms = orc.nifs_prove(prob, orc.transcript(wl["ring"]))
        orc.nifs_verify(prob, orc.transcript(wl["ring"]), proof)
        out["cases"][bench_case_key(config, log_w)] = dict(step_digests(proof, lc, f), oracle_step_ms=round(ms, 1), oracle_threads=orc.threads())
        print(bench_case_key(config, log_w), "step %.0f ms, total %.0f s" % (ms, time.time() - t0), flush=True)
        json.dump(out, open(BENCH_GOLDEN_PATH, "w"), indent=1)
