"""Inputs of the LatticeFold+ tests (set check / range check), shared by the oracle tests and the GPU parity tests.

The cases are the reference's own (crates/latticefold-plus/src/setchk.rs:358-495, rgchk.rs:344-433) plus seeded random ones.
Ring elements are coefficient-form arrays of d = 16 uint64 (the Frog ring `frog_ring::RqPoly`)."""
import numpy as np

RING_FROG = 2
D = 16
P_FROG = 15912092521325583641


def monomial(e, d=D):
    v = np.zeros(d, dtype=np.uint64)
    v[e] = 1
    return v


def one_plus_x(d=D):
    v = np.zeros(d, dtype=np.uint64)
    v[0] = v[1] = 1
    return v


def csr_from_entries(nrows, ncols, entries, d=D):
    """entries: {(row, col): d-vector}."""
    keys = sorted(entries)
    row_ptr = np.zeros(nrows + 1, dtype=np.uint64)
    for (r, _) in keys:
        row_ptr[r + 1] += 1
    row_ptr = np.cumsum(row_ptr).astype(np.uint64)
    col = np.array([c for (_, c) in keys], dtype=np.uint64)
    val = np.ascontiguousarray(np.stack([entries[k] for k in keys]).astype(np.uint64)) if keys else np.zeros((0, d), dtype=np.uint64)
    return dict(nrows=nrows, ncols=ncols, row_ptr=np.ascontiguousarray(row_ptr), col=np.ascontiguousarray(col), val=val)


def identity(n, d=D):      # SparseMatrix::identity
    return csr_from_entries(n, n, {(i, i): monomial(0, d) for i in range(n)}, d)


def random_monomial_matrix(nrows, ncols, seed, d=D, density=1.0):
    rng = np.random.default_rng(seed)
    ent = {}
    for r in range(nrows):
        for c in range(ncols):
            if rng.random() < density:
                ent[(r, c)] = monomial(int(rng.integers(0, d)), d)
    return csr_from_entries(nrows, ncols, ent, d)


def random_ring_sparse(nrows, ncols, per_row, seed, p=P_FROG, d=D, constant=False):
    """a sparse matrix of ring elements (the `M` argument of set_check / range_check).  constant=True: entries in Fq, as the
    range check's psi tests need (ct(psi * m * exp(a)) = m a only for m in Fq; rgchk.rs:385-387 uses such a matrix)."""
    rng = np.random.default_rng(seed)
    ent = {}
    for r in range(nrows):
        for c in rng.choice(ncols, size=min(per_row, ncols), replace=False):
            ent[(r, int(c))] = rng.integers(0, p, size=d, dtype=np.uint64)
            if constant:
                ent[(r, int(c))][1:] = 0
    return csr_from_entries(nrows, ncols, ent, d)


def random_monomial_vector(n, seed, d=D):
    rng = np.random.default_rng(seed)
    v = np.zeros((n, d), dtype=np.uint64)
    v[np.arange(n), rng.integers(0, d, size=n)] = 1
    return v


def set_check_cases():
    """name -> (nvars, sets, M, expect_accept); the first six are the reference's tests."""
    n = 4
    bad = identity(n)
    bad["val"] = bad["val"].copy()
    bad["val"][0] = one_plus_x()
    ones = np.zeros((n, D), dtype=np.uint64)
    ones[:, 0] = 1
    x2 = np.tile(monomial(2), (n, 1))
    badvec = ones.copy()
    badvec[0] = one_plus_x()
    cases = {
        "test_set_check": (2, [("matrix", identity(n))], [], True),
        "test_set_check_bad": (2, [("matrix", bad)], [], False),
        "test_set_check_batched": (2, [("matrix", identity(n)), ("matrix", identity(n))], [], True),
        "test_set_check_batched_bad": (2, [("matrix", identity(n)), ("matrix", bad)], [], False),
        "test_set_check_mix": (2, [("matrix", identity(n)), ("matrix", identity(n)), ("vector", ones), ("vector", x2)], [], True),
        "test_set_check_mix_bad": (2, [("matrix", identity(n)), ("matrix", identity(n)), ("vector", badvec)], [], False),
    }
    # seeded cases: rectangular monomial matrices (the range check's n x d shape), sparse columns, extra matrices M
    N = 64
    cases["rect_2sets"] = (6, [("matrix", random_monomial_matrix(N, 16, 1)), ("matrix", random_monomial_matrix(N, 16, 2)), ("vector", random_monomial_vector(N, 3))], [], True)
    cases["rect_with_M"] = (6, [("matrix", random_monomial_matrix(N, 16, 4, density=0.6)), ("matrix", random_monomial_matrix(N, 16, 5)), ("vector", random_monomial_vector(N, 6))],
                            [random_ring_sparse(N, N, 3, 7), identity(N)], True)
    cases["identity_256"] = (8, [("matrix", identity(256))], [], True)      # benches/setchk.rs SET_SIZES[0]
    short = random_monomial_matrix(48, 8, 8)      # fewer rows than 2^nvars: the tail of every MLE is zero
    cases["short_rows"] = (6, [("matrix", short), ("matrix", random_monomial_matrix(48, 8, 9))], [random_ring_sparse(48, 48, 2, 10)], True)
    return cases


def range_check_inputs(n, kappa, seed, L=1, p=P_FROG, d=D, k=2):
    """witnesses with coefficients inside the base-(d/2) k-digit balanced range, a random Ajtai matrix in coefficient form."""
    rng = np.random.default_rng(seed)
    bound = ((d // 2) ** k) // 2 - 1
    fs = np.zeros((L, n, d), dtype=np.uint64)
    for l in range(L):
        sm = rng.integers(-bound, bound + 1, size=(n, d))
        sm[rng.random(size=(n, d)) < 0.5] = 0
        fs[l] = np.where(sm < 0, p - (-sm).astype(np.uint64), sm.astype(np.uint64)).astype(np.uint64)
    A = rng.integers(0, p, size=(kappa, n, d), dtype=np.uint64)
    return np.ascontiguousarray(fs), np.ascontiguousarray(A)


def reference_range_check_f(n, d=D):      # rgchk.rs:354-361: f = [2 + 5X, 4 + X^2, 0, ...]
    f = np.zeros((1, n, d), dtype=np.uint64)
    f[0, 0, 0], f[0, 0, 1], f[0, 1, 0], f[0, 1, 2] = 2, 5, 4, 1
    return f


def frog_l(d=D, p=P_FROG):      # rgchk.rs:367-369: ceil(ln q / ln(d/2))
    import math
    return math.ceil(math.log(float(p)) / math.log(d / 2))


def r1cs_instance(n, seed, p=P_FROG, d=D):
    """A satisfied R1CS over the committed witness (r1cs.rs / plus.rs tests use decomposed identity systems): A = B = C = a 0/1 selection
    matrix with one entry scaled by 2 in A and C, and a 0/1 witness, so that (A f) * (B f) = C f holds as ring elements."""
    rng = np.random.default_rng(seed)
    f = np.zeros((n, d), dtype=np.uint64)
    f[:, 0] = rng.integers(0, 2, size=n).astype(np.uint64)
    f[0, 0] = 1
    perm = rng.permutation(n)
    ent = {(i, int(perm[i])): monomial(0, d) for i in range(n)}
    A = csr_from_entries(n, n, ent, d)
    two = dict(ent); two[(0, int(perm[0]))] = monomial(0, d) * np.uint64(2)
    A2 = csr_from_entries(n, n, two, d)
    return [A2, A, A2], f


class OraclePlus:
    """the flow of plus.rs (PlusProver::prove / PlusVerifier::verify) on the CPU oracle's entry points"""

    def __init__(self, orc, A, M, b, k, l, B, seed=None):
        self.o, self.A, self.M, self.b, self.k, self.l, self.B = orc, A, list(M), b, k, l, B
        self.tp, self.tv, self.acc = orc.plus_transcript(RING_FROG, seed), orc.plus_transcript(RING_FROG, seed), []

    def prove(self, comps):      # comps: list of (abc, f)
        lproof = []
        for abc, f in comps:
            linb, lp = self.o.plus_r1cs_linearize(RING_FROG, abc, f, self.tp)
            lproof.append(lp); self.acc.append(linb["f"])
        cmproof, x, g = self.o.plus_mlin_t(RING_FROG, np.stack(self.acc), self.A, self.b, self.k, self.l, self.M, self.tp)
        dproof, F = self.o.plus_decompose(RING_FROG, g, x["ro"], self.A, self.B, self.M)
        self.acc = [F[0], F[1]]
        return dict(linb2x=x, lproof=lproof, cmproof=cmproof, dproof=dproof)

    def verify(self, proof):
        ok = all(self.o.plus_r1cs_linearize_verify(RING_FROG, lp, self.tv) for lp in proof["lproof"])
        ok = ok and self.o.plus_cm_verify_t(RING_FROG, proof["cmproof"], len(self.M), self.tv)
        return ok and self.o.plus_decompose_verify(RING_FROG, proof["dproof"], self.A.shape[0], len(self.M), proof["linb2x"]["cm_g"], proof["linb2x"]["vo"], self.B)
