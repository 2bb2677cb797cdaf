"""Deterministic synthetic inputs for the LatticeFold prover step, shaped after the reference's bench generators.

  dummy R1CS (A = B = C = "identity" with one non-zero per row)      crates/latticefold/src/arith/r1cs.rs:155-223
  non-scalar variant (C = diag(z))                                    crates/latticefold/src/arith/r1cs.rs:170-223
  witnesses: all-ones / limbs 0..d-1 / uniform                        arith/r1cs.rs:279-306, build.rs:459
  CCS::from_r1cs_padded                                               crates/latticefold/src/arith/ccs.rs (arith.rs:122-171)
  accumulator = linearization of the same instance (e2e bench)        crates/latticefold/benches/utils.rs:640-677

PRNG is SplitMix64 (seed 0x4C46423230300000 + config id), base-field samples by rejection to [0, p), so the inputs do
not depend on any Rust RNG.  The Ajtai matrix holds kappa*n INDEPENDENT uniform ring elements (the reference's
`AjtaiCommitmentScheme::rand` repeats one sample, commitment_scheme.rs:30-32).

Ring arithmetic needed to derive `f` from `w_ccs`, the commitment and the accumulator is delegated to an `ops`
object (the GPU library in bench.py, the CPU oracle in CPU-only tests); nothing here computes on ring elements.
"""
import numpy as np

RING_GOLDILOCKS, RING_BABYBEAR, RING_FROG = 0, 1, 2
RINGS = {
    RING_GOLDILOCKS: dict(name="goldilocks", p=0xFFFFFFFF00000001, d=24, S=8, tau=3),
    RING_BABYBEAR: dict(name="babybear", p=2013265921, d=72, S=8, tau=9),
    RING_FROG: dict(name="frog", p=15912092521325583641, d=16, S=4, tau=4),
}
SEED_BASE = 0x4C46423230300000
MASK = (1 << 64) - 1


def splitmix64(seed, count):
    """count outputs of SplitMix64 started at `seed` (vectorised)."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, count + 1, dtype=np.uint64)
        z = np.uint64(seed & MASK) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform_field(p, count, seed):
    """`count` uniform samples in [0, p) by rejection on SplitMix64 outputs masked to bitlen(p)."""
    bits = int(p).bit_length()
    mask = np.uint64((1 << bits) - 1)
    out = np.empty(count, dtype=np.uint64)
    have, s = 0, seed & MASK
    while have < count:
        need = count - have
        draw = splitmix64(s, need + need // 8 + 16) & mask
        s = (s + 0x632BE59BD9B4E019 * (len(draw) + 1)) & MASK
        good = draw[draw < np.uint64(p)][:need]
        out[have:have + len(good)] = good
        have += len(good)
    return out


def one(ring, count=1):
    R = RINGS[ring]
    o = np.zeros((count, R["d"]), dtype=np.uint64)
    o[:, ::R["tau"]] = 1
    return o


def dummy_csr(ring, rows_used, ncols, m, diag=None):
    """rows_used rows with one non-zero (value one, or diag[i]) at column i; padded with empty rows to m."""
    R = RINGS[ring]
    row_ptr = np.minimum(np.arange(m + 1, dtype=np.uint64), np.uint64(rows_used))
    col = np.arange(rows_used, dtype=np.uint64)
    val = one(ring, rows_used) if diag is None else np.ascontiguousarray(diag[:rows_used], dtype=np.uint64)
    return dict(nrows=m, ncols=ncols, row_ptr=np.ascontiguousarray(row_ptr), col=col, val=val)


def make_witness(ring, W, kind, seed):
    R = RINGS[ring]
    if kind == "scalar":
        return one(ring, W)
    if kind == "non_scalar":
        return np.ascontiguousarray(np.tile(np.arange(R["d"], dtype=np.uint64), (W, 1)))
    if kind == "uniform":
        return uniform_field(R["p"], W * R["d"], seed ^ 0x77).reshape(W, R["d"])
    raise ValueError(kind)


def sf_mul(ring, a, b, nu):
    """slot-wise product of NTT-form ring elements (count x d) in Fq[Y]/(Y^tau - nu), with Python integers (test-sized inputs only)"""
    R = RINGS[ring]; p, S, tau = R["p"], R["S"], R["tau"]
    x = a.reshape(-1, S, tau).astype(object); y = b.reshape(-1, S, tau).astype(object)
    out = np.zeros_like(x)
    for i in range(tau):
        for j in range(tau):
            t = x[:, :, i] * y[:, :, j]
            if i + j < tau:
                out[:, :, i + j] = (out[:, :, i + j] + t) % p
            else:
                out[:, :, i + j - tau] = (out[:, :, i + j - tau] + nu * t) % p
    return np.ascontiguousarray(out.astype(np.uint64).reshape(-1, R["d"]))


def make_ccs(ring, W, L, kind, w_ccs, x_len=1, degree=2, ops=None):
    """The reference's dummy R1CS as a padded CCS (degree 2), or its dummy degree-three CCS (arith/ccs.rs:14-43: A = B = C =
    identity-like, D = diag(z^2), (Az)(Bz)(Cz) - Dz = 0).  z = x || 1 || w with x = ones."""
    R = RINGS[ring]
    ncols = x_len + W + 1
    rows = x_len + W + 1
    n = W * L
    m = 1
    while m < max((ncols - 1 - 1) * L, n):
        m <<= 1
    z = np.concatenate([one(ring, x_len), one(ring, 1), w_ccs])
    if degree == 3:
        neg_one = np.zeros((1, R["d"]), dtype=np.uint64)
        neg_one[:, ::R["tau"]] = R["p"] - 1
        Ms = [dummy_csr(ring, rows, ncols, m) for _ in range(3)] + [dummy_csr(ring, rows, ncols, m, diag=ops.ntt_mul(ring, z, z))]
        return dict(m=m, n_ccs=ncols, l=1, t=4, q=2, d=3, s=m.bit_length() - 1, M=Ms, S=[[0, 1, 2], [3]],
                    c=np.ascontiguousarray(np.concatenate([one(ring, 1), neg_one])))
    A = dummy_csr(ring, rows, ncols, m)
    Bm = dummy_csr(ring, rows, ncols, m)
    Cm = dummy_csr(ring, rows, ncols, m, diag=None if kind == "scalar" else z)
    neg_one = np.zeros((1, R["d"]), dtype=np.uint64)
    neg_one[:, ::R["tau"]] = R["p"] - 1
    s = m.bit_length() - 1
    return dict(m=m, n_ccs=ncols, l=1, t=3, q=2, d=2, s=s, M=[A, Bm, Cm], S=[[0, 1], [2]],
                c=np.ascontiguousarray(np.concatenate([one(ring, 1), neg_one])))


def make_instance(ring, W, B, L, b, K, kappa, kind="non_scalar", config_id=0, ops=None, with_acc=True, x_len=1, degree=2):
    """One prover-step input set.  `ops` must offer witness_f_from_w_ccs(ring, w_ccs, B, L) -> f (n x d, NTT form),
    commit(ring, A, f) -> (kappa x d) and linearize(problem) -> dict(r, v, cm, u, x_w, h)."""
    R = RINGS[ring]
    seed = (SEED_BASE + config_id) & MASK
    n = W * L
    w_ccs = make_witness(ring, W, kind, seed)
    ccs = make_ccs(ring, W, L, kind, w_ccs, x_len, degree, ops)
    A = uniform_field(R["p"], kappa * n * R["d"], seed).reshape(kappa, n, R["d"])
    prob = dict(ring=ring, B=B, L=L, b=b, K=K, kappa=kappa, n=n, W=W, A=A, ccs=ccs, w_ccs=w_ccs,
                cm_i_x_ccs=one(ring, x_len), constraints=x_len + W + 1, kind=kind)
    if ops is not None:
        f = np.ascontiguousarray(ops.witness_f_from_w_ccs(ring, w_ccs, B, L))
        prob["w_i_f"] = f
        prob["cm_i_cm"] = np.ascontiguousarray(ops.commit(ring, A, f))
        if with_acc:
            prob["acc"] = ops.linearize(prob)
            prob["w_acc_f"] = f
    return prob


def split_lcccs(ring, prob, words):
    """Unpack the flat LCCCS serialisation (r, v, cm, u, x_w, h) used on the C boundary."""
    R = RINGS[ring]
    d, tau = R["d"], R["tau"]
    ccs = prob["ccs"]
    sizes = [("r", ccs["s"]), ("v", tau), ("cm", prob["kappa"]), ("u", ccs["t"]), ("x_w", ccs["l"]), ("h", 1)]
    out, o = {}, 0
    for name, cnt in sizes:
        out[name] = np.ascontiguousarray(words[o:o + cnt * d].reshape(cnt, d))
        o += cnt * d
    assert o == words.size
    return out


# ---------------------------------------------------------------------------------------------------------------
# The workloads bench.py times (BASELINE.json configs) -- shared by bench.py, the golden-digest
# generator (tools/make_bench_golden.py) and the full-size GPU tests, so that all three prove the same instance.
def bench_workload(config, log_w):
    """c2: configs[1] = Goldilocks, DP (B, L, b, K) = (65536, 4, 2, 16), kappa = 26 (benches/config.toml goldilocks row n=32768
    extrapolated, as SURVEY.md 8 does), dummy R1CS, non-scalar witness; weak-scaled over G ranks it is configs[3] (C4) at log_w = 19.
    c3: configs[2] = BabyBear ring, BabyBearDP (256, 4, 2, 8) (decomposition_parameters.rs:98-104), kappa = 8, the reference's
    degree-three CCS (arith/ccs.rs:14-43)."""
    if config == "c2":
        return dict(config="c2", ring=RING_GOLDILOCKS, W=1 << log_w, B=1 << 16, L=4, b=2, K=16, kappa=26, kind="non_scalar", degree=2)
    if config == "c3":
        return dict(config="c3", ring=RING_BABYBEAR, W=1 << log_w, B=1 << 8, L=4, b=2, K=8, kappa=8, kind="non_scalar", degree=3)
    raise ValueError(config)


def bench_instance(wl, rank, world, ops=None):
    """Synthetic inputs of one step, holding only this rank's column slice of the Ajtai matrix (kappa x n/world independent
    uniform ring elements, SplitMix64 seeded per rank) -- the full 8-GPU matrix would be 10.5 GB per process."""
    ring = wl["ring"]; R = RINGS[ring]
    seed = (SEED_BASE + 100) & MASK
    n = wl["W"] * wl["L"]
    w_ccs = make_witness(ring, wl["W"], wl["kind"], seed)
    ccs = make_ccs(ring, wl["W"], wl["L"], wl["kind"], w_ccs, 1, wl.get("degree", 2), ops)
    row = (n // world) * R["d"]
    if wl["kappa"] * row <= 1 << 28:
        A = uniform_field(R["p"], wl["kappa"] * row, seed + 7919 * (rank + 1)).reshape(wl["kappa"], n // world, R["d"])
    else:       # large matrices row by row (bounded temporaries: configs[2] at 2^20 is 19 GB of host limbs): row i seeded on its own
        A = np.empty((wl["kappa"], n // world, R["d"]), dtype=np.uint64)
        for i in range(wl["kappa"]):
            A[i] = uniform_field(R["p"], row, seed + 7919 * (rank + 1) + 104729 * (i + 1)).reshape(n // world, R["d"])
    return dict(ring=ring, B=wl["B"], L=wl["L"], b=wl["b"], K=wl["K"], kappa=wl["kappa"], n=n, W=wl["W"], A=A, ccs=ccs, w_ccs=w_ccs,
                cm_i_x_ccs=one(ring, 1), constraints=1 + wl["W"] + 1, kind=wl["kind"])
