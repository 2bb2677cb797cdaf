"""Shared test helpers (CPU oracle adapters, random ring data)."""
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
