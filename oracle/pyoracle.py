"""TEST INFRASTRUCTURE ONLY: ctypes front-end of the CPU oracle (oracle/_build/liblfo.so).

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the
product package.  Ring elements are numpy uint64 arrays whose last axis is d canonical limbs.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liblfo.so")

u64p = C.POINTER(C.c_uint64)
i32p = C.POINTER(C.c_int)


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("capi.cpp", "ntt.hpp", "ring.hpp", "transcript.hpp", "sumcheck.hpp", "protocol.hpp", "lfplus.hpp")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "-s"], stdout=subprocess.DEVNULL)
    return LIB_PATH


class Csr(C.Structure):
    _fields_ = [("nrows", C.c_uint64), ("ncols", C.c_uint64), ("row_ptr", u64p), ("col", u64p), ("val", u64p)]


class PlusSet(C.Structure):
    """One monomial set of the LatticeFold+ set check (kind 0: sparse matrix, 1: vector); layout of `lfo_plus_set` / `lf_plus_set`."""
    _fields_ = [("kind", C.c_int32), ("pad", C.c_int32), ("m", Csr), ("v", u64p), ("n", C.c_uint64)]


def make_csr_array(mats, csr_cls=Csr):
    arr = (csr_cls * max(len(mats), 1))()
    for j, M in enumerate(mats):
        arr[j].nrows, arr[j].ncols = M["nrows"], M["ncols"]
        arr[j].row_ptr, arr[j].col, arr[j].val = ptr(M["row_ptr"]), ptr(M["col"]), ptr(M["val"])
    return arr


def make_plus_sets(sets, set_cls=None, csr_cls=Csr):
    """sets: list of ("matrix", csr dict) / ("vector", n x d uint64 array)."""
    set_cls = set_cls or PlusSet
    arr = (set_cls * max(len(sets), 1))()
    for i, (kind, x) in enumerate(sets):
        if kind == "matrix":
            arr[i].kind = 0
            arr[i].m.nrows, arr[i].m.ncols = x["nrows"], x["ncols"]
            arr[i].m.row_ptr, arr[i].m.col, arr[i].m.val = ptr(x["row_ptr"]), ptr(x["col"]), ptr(x["val"])
        else:
            arr[i].kind = 1
            arr[i].v, arr[i].n = ptr(x), x.shape[0]
    return arr


class Problem(C.Structure):
    """Flat description of one NIFS prover step; field-for-field the layout of `lfo_problem` (oracle/capi.cpp)
    and of `lf_problem` (include/lf_b200.h)."""
    _fields_ = [("ring", C.c_int), ("L", C.c_int), ("K", C.c_int), ("B_lo", C.c_uint64), ("B_hi", C.c_uint64), ("b", C.c_uint64),
                ("kappa", C.c_uint64), ("n", C.c_uint64), ("A", u64p),
                ("m", C.c_uint64), ("n_ccs", C.c_uint64), ("l", C.c_uint64), ("t", C.c_uint64), ("q", C.c_uint64), ("d", C.c_uint64), ("s", C.c_uint64),
                ("M", C.POINTER(Csr)), ("S_flat", i32p), ("S_len", i32p), ("c", u64p),
                ("acc_r", u64p), ("acc_v", u64p), ("acc_cm", u64p), ("acc_u", u64p), ("acc_x_w", u64p), ("acc_h", u64p),
                ("w_acc_f", u64p), ("cm_i_cm", u64p), ("cm_i_x_ccs", u64p), ("w_i_f", u64p)]


def ptr(a):
    if a is None:
        return None
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(u64p)


def make_problem(p, problem_cls=Problem, csr_cls=Csr):
    """p: dict produced by latticefold_b200.synth.make_instance (numpy arrays).  Returns (struct, keepalive)."""
    keep = []
    P = problem_cls()
    P.ring, P.L, P.K = p["ring"], p["L"], p["K"]
    P.B_lo, P.B_hi, P.b = p["B"] & (2**64 - 1), p["B"] >> 64, p["b"]
    P.kappa, P.n = p["kappa"], p["n"]
    P.A = ptr(p.get("A"))
    ccs = p["ccs"]
    for k in ("m", "n_ccs", "l", "t", "q", "d", "s"):
        setattr(P, k, ccs[k])
    arr = (csr_cls * ccs["t"])()
    for j, M in enumerate(ccs["M"]):
        arr[j].nrows, arr[j].ncols = M["nrows"], M["ncols"]
        arr[j].row_ptr, arr[j].col, arr[j].val = ptr(M["row_ptr"]), ptr(M["col"]), ptr(M["val"])
    keep.append(arr)
    P.M = arr
    S_flat = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=np.int32) for s in ccs["S"]]))
    S_len = np.ascontiguousarray(np.array([len(s) for s in ccs["S"]], dtype=np.int32))
    keep += [S_flat, S_len]
    P.S_flat, P.S_len = S_flat.ctypes.data_as(i32p), S_len.ctypes.data_as(i32p)
    P.c = ptr(ccs["c"])
    acc = p.get("acc")
    if acc is not None:
        P.acc_r, P.acc_v, P.acc_cm, P.acc_u, P.acc_x_w, P.acc_h = (ptr(acc[k]) for k in ("r", "v", "cm", "u", "x_w", "h"))
    P.w_acc_f = ptr(p.get("w_acc_f"))
    P.cm_i_cm, P.cm_i_x_ccs = ptr(p.get("cm_i_cm")), ptr(p["cm_i_x_ccs"])
    P.w_i_f = ptr(p.get("w_i_f"))
    keep.append(p)
    return P, keep


class Oracle:
    def __init__(self):
        build()
        L = self.lib = C.CDLL(LIB_PATH)
        L.lfo_last_error.restype = C.c_char_p
        L.lfo_tr_new.restype = C.c_void_p
        L.lfo_tr_clone.restype = C.c_void_p
        L.lfo_tr_clone.argtypes = [C.c_void_p]
        L.lfo_tr_free.argtypes = [C.c_void_p]
        for f in ("lfo_tr_absorb", "lfo_tr_absorb_base", "lfo_tr_squeeze_base"):
            getattr(L, f).argtypes = [C.c_void_p, u64p, C.c_size_t]
        L.lfo_tr_absorb_tag.argtypes = [C.c_void_p, C.c_char_p]
        L.lfo_tr_absorb_u64.argtypes = [C.c_void_p, C.c_uint64]
        L.lfo_tr_get_challenge.argtypes = [C.c_void_p, u64p]
        L.lfo_tr_get_short_challenge.argtypes = [C.c_void_p, u64p]
        L.lfo_tr_state.argtypes = [C.c_void_p, u64p]
        L.lfo_proof_words.restype = C.c_uint64
        L.lfo_lcccs_words.restype = C.c_uint64
        L.lfo_crt.argtypes = L.lfo_icrt.argtypes = [C.c_int, u64p, u64p, C.c_size_t]
        L.lfo_coeff_mul.argtypes = [C.c_int, u64p, u64p, u64p]
        L.lfo_ntt_mul.argtypes = [C.c_int, u64p, u64p, u64p, C.c_size_t]
        L.lfo_gadget_decompose.argtypes = [C.c_int, u64p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_int, u64p]
        L.lfo_gadget_recompose.argtypes = [C.c_int, u64p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_int, u64p]
        L.lfo_decompose_to_vec.argtypes = [C.c_int, u64p, C.c_size_t, C.c_uint64, C.c_int, u64p]
        L.lfo_fhat.argtypes = [C.c_int, u64p, C.c_size_t, u64p, u64p]
        L.lfo_commit.argtypes = [C.c_int, u64p, C.c_size_t, C.c_size_t, u64p, C.c_size_t, u64p]
        L.lfo_spmv.argtypes = [C.c_int, C.c_size_t, C.c_size_t, u64p, u64p, u64p, u64p, C.c_size_t, u64p]
        L.lfo_eq_table.argtypes = [C.c_int, u64p, C.c_int, u64p]
        L.lfo_eq_eval.argtypes = [C.c_int, u64p, u64p, C.c_int, u64p]
        L.lfo_evaluate_mles.argtypes = [C.c_int, u64p, C.c_int, C.c_size_t, C.c_int, u64p, C.c_int, u64p]
        L.lfo_rot_lin_combination.argtypes = [C.c_int, u64p, u64p, C.c_int, u64p]
        L.lfo_ntt_root.restype = C.c_uint64
        L.lfo_ntt_root.argtypes = [C.c_int, C.c_int]
        L.lfo_ntt_naive.argtypes = [C.c_int, C.c_int, u64p, u64p, C.c_int]
        L.lfo_ntt_fast.argtypes = [C.c_int, C.c_int, u64p, u64p, C.c_size_t, C.c_int]
        L.lfo_ntt_schoolbook.argtypes = [C.c_int, C.c_int, u64p, u64p, u64p]
        L.lfo_short_challenge_from_bytes.argtypes = [C.c_int, C.POINTER(C.c_uint8), u64p]
        L.lfo_sumcheck_prove.argtypes = [C.c_int, C.c_void_p, u64p, C.c_int, C.c_size_t, u64p, C.c_int, C.c_int, C.c_int,
                                         C.c_int, u64p, i32p, i32p, C.c_int, C.c_int, u64p, u64p, u64p, u64p]
        L.lfo_sumcheck_verify.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, u64p, u64p, u64p, u64p]
        L.lfo_linearize.argtypes = [C.POINTER(Problem), C.c_void_p, u64p, u64p]
        L.lfo_linearization_verify.argtypes = [C.POINTER(Problem), C.c_void_p, u64p, u64p]
        L.lfo_nifs_prove.argtypes = [C.POINTER(Problem), C.c_void_p, u64p, u64p, u64p, C.POINTER(C.c_double)]
        L.lfo_nifs_verify.argtypes = [C.POINTER(Problem), C.c_void_p, u64p, u64p]
        L.lfo_proof_words.argtypes = [C.POINTER(Problem)]
        L.lfo_lcccs_words.argtypes = [C.POINTER(Problem)]
        self._info = {}

    # ---------------------------------------------------------------- helpers
    def err(self):
        return self.lib.lfo_last_error().decode()

    def check(self, rc):
        if rc != 0:
            raise OracleError(rc, self.err())

    def info(self, ring):
        if ring not in self._info:
            o = np.zeros(8, dtype=np.uint64)
            self.check(self.lib.lfo_ring_info(ring, ptr(o)))
            self._info[ring] = dict(p=int(o[0]), d=int(o[1]), S=int(o[2]), tau=int(o[3]), g=int(o[4]), nu=int(o[5]), trinomial=bool(o[6]), cs_bytes=int(o[7]))
        return self._info[ring]

    # ---- negacyclic NTT over Z_p[X]/(X^N + 1) (oracle/ntt.hpp); field 0 = Goldilocks, 1 = BabyBear; values as uint64
    NTT_P = {0: 0xFFFFFFFF00000001, 1: 2013265921}

    def ntt_root(self, field, log_n):
        return int(self.lib.lfo_ntt_root(field, log_n))

    def ntt_naive(self, field, log_n, a, inverse=False):
        a = np.ascontiguousarray(a, dtype=np.uint64); o = np.empty_like(a)
        self.check(self.lib.lfo_ntt_naive(field, log_n, ptr(a), ptr(o), int(inverse))); return o

    def ntt(self, field, log_n, a, inverse=False):
        """a: (batch, N) uint64 -> fast CPU transform of every row (textbook radix-2, OpenMP over the batch)"""
        a = np.ascontiguousarray(a, dtype=np.uint64); o = np.empty_like(a)
        self.check(self.lib.lfo_ntt_fast(field, log_n, ptr(a), ptr(o), a.size >> log_n, int(inverse))); return o

    def ntt_schoolbook(self, field, log_n, a, b):
        a = np.ascontiguousarray(a, dtype=np.uint64); b = np.ascontiguousarray(b, dtype=np.uint64); o = np.empty_like(a)
        self.check(self.lib.lfo_ntt_schoolbook(field, log_n, ptr(a), ptr(b), ptr(o))); return o

    def threads(self):
        return self.lib.lfo_num_threads()

    def set_threads(self, n):
        self.lib.lfo_set_num_threads(n)

    # ---------------------------------------------------------------- ring ops
    def crt(self, ring, a):
        a = np.ascontiguousarray(a, dtype=np.uint64); o = np.empty_like(a)
        self.check(self.lib.lfo_crt(ring, ptr(a), ptr(o), a.size // self.info(ring)["d"])); return o

    def icrt(self, ring, a):
        a = np.ascontiguousarray(a, dtype=np.uint64); o = np.empty_like(a)
        self.check(self.lib.lfo_icrt(ring, ptr(a), ptr(o), a.size // self.info(ring)["d"])); return o

    def coeff_mul(self, ring, a, b):
        o = np.empty_like(a); self.check(self.lib.lfo_coeff_mul(ring, ptr(a), ptr(b), ptr(o))); return o

    def ntt_mul(self, ring, a, b):
        o = np.empty_like(a); self.check(self.lib.lfo_ntt_mul(ring, ptr(a), ptr(b), ptr(o), a.size // self.info(ring)["d"])); return o

    def gadget_decompose(self, ring, a, B, L):
        d = self.info(ring)["d"]; n = a.size // d; o = np.empty((n * L, d), dtype=np.uint64)
        self.check(self.lib.lfo_gadget_decompose(ring, ptr(a), n, B & (2**64 - 1), B >> 64, L, ptr(o))); return o

    def gadget_recompose(self, ring, a, B, L):
        d = self.info(ring)["d"]; n = a.size // d; o = np.empty((n // L, d), dtype=np.uint64)
        self.check(self.lib.lfo_gadget_recompose(ring, ptr(a), n, B & (2**64 - 1), B >> 64, L, ptr(o))); return o

    def decompose_to_vec(self, ring, a, b, K):
        d = self.info(ring)["d"]; n = a.size // d; o = np.empty((K, n, d), dtype=np.uint64)
        self.check(self.lib.lfo_decompose_to_vec(ring, ptr(a), n, b, K, ptr(o))); return o

    def fhat(self, ring, f_coeff):
        i = self.info(ring); n = f_coeff.size // i["d"]; o = np.empty((i["tau"], n, i["d"]), dtype=np.uint64); lens = np.zeros(i["tau"], dtype=np.uint64)
        self.check(self.lib.lfo_fhat(ring, ptr(f_coeff), n, ptr(o), ptr(lens))); return o, lens

    def commit(self, ring, A, f):
        d = self.info(ring)["d"]; kappa, n = A.shape[0], A.shape[1]; o = np.empty((kappa, d), dtype=np.uint64)
        self.check(self.lib.lfo_commit(ring, ptr(A), kappa, n, ptr(f), f.size // d, ptr(o))); return o

    def spmv(self, ring, M, z):
        d = self.info(ring)["d"]; o = np.empty((M["nrows"], d), dtype=np.uint64)
        self.check(self.lib.lfo_spmv(ring, M["nrows"], M["ncols"], ptr(M["row_ptr"]), ptr(M["col"]), ptr(M["val"]), ptr(z), z.size // d, ptr(o))); return o

    def eq_table(self, ring, r):
        d = self.info(ring)["d"]; s = r.size // d; o = np.empty((1 << s, d), dtype=np.uint64)
        self.check(self.lib.lfo_eq_table(ring, ptr(r), s, ptr(o))); return o

    def eq_eval(self, ring, x, y):
        d = self.info(ring)["d"]; o = np.empty(d, dtype=np.uint64)
        self.check(self.lib.lfo_eq_eval(ring, ptr(x), ptr(y), x.size // d, ptr(o))); return o

    def evaluate_mles(self, ring, mles, nv, point):
        d = self.info(ring)["d"]; count, ln = mles.shape[0], mles.shape[1]; o = np.empty((count, d), dtype=np.uint64)
        self.check(self.lib.lfo_evaluate_mles(ring, ptr(mles), count, ln, nv, ptr(point), point.size // d, ptr(o))); return o

    def rot_lin_combination(self, ring, rho_coeff, theta):
        i = self.info(ring); o = np.empty((i["tau"], i["d"]), dtype=np.uint64)
        self.check(self.lib.lfo_rot_lin_combination(ring, ptr(rho_coeff), ptr(theta), rho_coeff.shape[0], ptr(o))); return o

    def short_challenge_from_bytes(self, ring, bs):
        d = self.info(ring)["d"]; o = np.empty(d, dtype=np.uint64); b = (C.c_uint8 * len(bs))(*bs)
        self.check(self.lib.lfo_short_challenge_from_bytes(ring, b, ptr(o))); return o

    # ---------------------------------------------------------------- transcript
    def transcript(self, ring):
        return OracleTranscript(self, ring)

    # ---------------------------------------------------------------- sumcheck
    def sumcheck_prove(self, ring, tr, mles, nv, degree, comb, lens=None, want_final=False):
        """mles: (M, len, d).  comb: dict(kind='products'|'lin', coef=(nterms,d), idx=[[...]]) or dict(kind='fold', mu=(n_mu,d), b=int)."""
        i = self.info(ring); d, tau = i["d"], i["tau"]; M, ln = mles.shape[0], mles.shape[1]
        msgs = np.empty((nv, degree + 1, d), dtype=np.uint64); point = np.empty((nv, tau), dtype=np.uint64)
        final = np.empty((M, d), dtype=np.uint64) if want_final else None
        lens_a = None if lens is None else np.ascontiguousarray(lens, dtype=np.uint64)
        if comb["kind"] == "fold":
            rc = self.lib.lfo_sumcheck_prove(ring, tr.h, ptr(mles), M, ln, ptr(lens_a), nv, degree, 2, 0, None, None, None,
                                             comb["mu"].shape[0], comb["b"], ptr(comb["mu"]), ptr(msgs), ptr(point), ptr(final))
        else:
            idx_flat = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.int32) for x in comb["idx"]]))
            idx_len = np.ascontiguousarray(np.array([len(x) for x in comb["idx"]], dtype=np.int32))
            rc = self.lib.lfo_sumcheck_prove(ring, tr.h, ptr(mles), M, ln, ptr(lens_a), nv, degree, 1 if comb["kind"] == "lin" else 0,
                                             len(comb["idx"]), ptr(comb["coef"]), idx_flat.ctypes.data_as(i32p), idx_len.ctypes.data_as(i32p),
                                             0, 0, None, ptr(msgs), ptr(point), ptr(final))
        self.check(rc)
        return (msgs, point, final) if want_final else (msgs, point)

    def sumcheck_verify(self, ring, tr, nv, degree, claimed_sum, msgs):
        i = self.info(ring); exp = np.empty(i["d"], dtype=np.uint64); point = np.empty((nv, i["tau"]), dtype=np.uint64)
        self.check(self.lib.lfo_sumcheck_verify(ring, tr.h, nv, degree, ptr(claimed_sum), ptr(np.ascontiguousarray(msgs)), ptr(exp), ptr(point)))
        return exp, point

    # ---------------------------------------------------------------- protocol
    def linearize(self, prob, tr):
        P, keep = make_problem(prob)
        d = self.info(prob["ring"])["d"]; tau = self.info(prob["ring"])["tau"]; ccs = prob["ccs"]
        lc = np.empty(self.lib.lfo_lcccs_words(C.byref(P)), dtype=np.uint64)
        pf = np.empty((ccs["s"] * (ccs["d"] + 2) + tau + ccs["t"]) * d, dtype=np.uint64)
        self.check(self.lib.lfo_linearize(C.byref(P), tr.h, ptr(lc), ptr(pf)))
        return lc, pf

    def linearization_verify(self, prob, tr, lin_proof):
        P, keep = make_problem(prob)
        lc = np.empty(self.lib.lfo_lcccs_words(C.byref(P)), dtype=np.uint64)
        self.check(self.lib.lfo_linearization_verify(C.byref(P), tr.h, ptr(np.ascontiguousarray(lin_proof)), ptr(lc)))
        return lc

    def nifs_prove(self, prob, tr, want_f=True):
        P, keep = make_problem(prob)
        d = self.info(prob["ring"])["d"]
        proof = np.empty(self.lib.lfo_proof_words(C.byref(P)), dtype=np.uint64)
        lc = np.empty(self.lib.lfo_lcccs_words(C.byref(P)), dtype=np.uint64)
        f = np.empty((prob["n"], d), dtype=np.uint64) if want_f else None
        ms = (C.c_double * 4)()
        self.check(self.lib.lfo_nifs_prove(C.byref(P), tr.h, ptr(proof), ptr(lc), ptr(f), ms))
        return proof, lc, f, ms[0]

    def nifs_verify(self, prob, tr, proof):
        P, keep = make_problem(prob)
        lc = np.empty(self.lib.lfo_lcccs_words(C.byref(P)), dtype=np.uint64)
        self.check(self.lib.lfo_nifs_verify(C.byref(P), tr.h, ptr(np.ascontiguousarray(proof)), ptr(lc)))
        return lc

    # ---- LatticeFold+ (oracle/lfplus.hpp): set check and range check on the coefficient-form ring; results are flat uint64 images
    def _plus_setup(self):
        L = self.lib
        if getattr(self, "_plus_ready", False):
            return
        L.lfo_plus_set_check.restype = C.c_long
        L.lfo_plus_set_check.argtypes = [C.c_int, C.c_int, C.POINTER(PlusSet), C.c_int, C.POINTER(Csr), C.c_int, u64p, C.c_size_t, u64p, C.c_size_t]
        L.lfo_plus_set_check_verify.argtypes = [C.c_int, u64p, C.c_size_t, u64p, C.c_size_t]
        L.lfo_plus_rg_from_f.argtypes = [C.c_int, u64p, C.c_size_t, u64p, C.c_size_t, C.c_uint64, C.c_int, C.c_int, u64p, u64p, u64p]
        L.lfo_plus_range_check.restype = C.c_long
        L.lfo_plus_range_check.argtypes = [C.c_int, C.c_int, C.c_int, u64p, C.c_size_t, u64p, C.c_size_t, C.c_uint64, C.c_int, C.c_int, C.POINTER(Csr), C.c_int, u64p, C.c_size_t, u64p, C.c_size_t]
        L.lfo_plus_range_check_verify.argtypes = [C.c_int, u64p, C.c_size_t, u64p, C.c_size_t]
        L.lfo_plus_cm_prove.restype = C.c_long
        L.lfo_plus_cm_prove.argtypes = [C.c_int, C.c_int, C.c_int, u64p, C.c_size_t, u64p, C.c_size_t, C.c_uint64, C.c_int, C.c_int, C.POINTER(Csr), C.c_int, u64p, C.c_size_t,
                                        u64p, C.c_size_t, u64p, C.c_size_t, u64p]
        L.lfo_plus_cm_verify.argtypes = [C.c_int, u64p, C.c_size_t, C.POINTER(Csr), C.c_int, u64p, C.c_size_t, u64p, C.c_size_t]
        L.lfo_plus_mlin.restype = C.c_long
        L.lfo_plus_mlin.argtypes = [C.c_int, C.c_int, u64p, C.c_size_t, u64p, C.c_size_t, C.c_uint64, C.c_int, C.c_int, C.POINTER(Csr), C.c_int, u64p, C.c_size_t, u64p, C.c_size_t, u64p, u64p]
        L.lfo_plus_decompose.argtypes = [C.c_int, u64p, C.c_size_t, u64p, C.POINTER(Csr), C.c_int, u64p, C.c_size_t, C.c_uint64, u64p, u64p]
        L.lfo_plus_decompose_verify.argtypes = [C.c_int, u64p, C.c_size_t, C.c_int, u64p, u64p, C.c_uint64]
        L.lfo_plus_r1cs_linearize.restype = C.c_long
        L.lfo_plus_r1cs_linearize.argtypes = [C.c_int, C.POINTER(Csr), u64p, C.c_size_t, C.c_void_p, u64p, C.c_size_t]
        L.lfo_plus_r1cs_linearize_verify.argtypes = [C.c_int, u64p, C.c_size_t, C.c_void_p]
        L.lfo_plus_tr_new.restype = C.c_void_p
        L.lfo_plus_tr_new.argtypes = [C.c_int, u64p, C.c_size_t]
        L.lfo_plus_tr_free.argtypes = [C.c_void_p]
        L.lfo_plus_tr_challenge.restype = C.c_uint64
        L.lfo_plus_tr_challenge.argtypes = [C.c_void_p]
        L.lfo_plus_mlin_t.restype = C.c_long
        L.lfo_plus_mlin_t.argtypes = [C.c_int, C.c_int, u64p, C.c_size_t, u64p, C.c_size_t, C.c_uint64, C.c_int, C.c_int, C.POINTER(Csr), C.c_int, C.c_void_p, u64p, C.c_size_t, u64p, u64p]
        L.lfo_plus_cm_verify_t.argtypes = [C.c_int, u64p, C.c_size_t, C.c_int, C.c_void_p]
        L.lfo_plus_mat_vec.argtypes = [C.c_int, u64p, C.c_size_t, C.c_size_t, u64p, u64p]
        L.lfo_plus_tensor.argtypes = [C.c_int, u64p, C.c_int, u64p]
        L.lfo_plus_ring_mul.argtypes = [C.c_int, u64p, u64p, u64p]
        self._plus_ready = True

    @staticmethod
    def _seed(seed):
        seed = np.ascontiguousarray(np.asarray(seed if seed is not None else [], dtype=np.uint64))
        return seed, (ptr(seed) if seed.size else None), seed.size

    def _plus_call(self, fn):
        cap = 1 << 16
        while True:
            out = np.zeros(cap, dtype=np.uint64)
            n = fn(out, cap)
            if n < 0:
                raise OracleError(int(n), self.err())
            if n <= cap:
                return out[:n].copy()
            cap = int(n)

    def plus_set_check(self, ring, nvars, sets, M=(), seed=None):
        self._plus_setup()
        sa, ma = make_plus_sets(sets), make_csr_array(list(M))
        sd, sp, sn = self._seed(seed)
        return self._plus_call(lambda out, cap: self.lib.lfo_plus_set_check(ring, nvars, sa, len(sets), ma, len(M), sp, sn, ptr(out), cap))

    def plus_set_check_verify(self, ring, words, seed=None):
        self._plus_setup()
        words = np.ascontiguousarray(words, dtype=np.uint64)
        sd, sp, sn = self._seed(seed)
        rc = self.lib.lfo_plus_set_check_verify(ring, ptr(words), words.size, sp, sn)
        if rc < 0:
            raise OracleError(rc, self.err())
        return bool(rc)

    def plus_rg_from_f(self, ring, f, A, b, k, l):
        """RgInstance::from_f: returns (tau[n], fcoms[3, kappa, d], comM_f[k, kappa, d, d])."""
        self._plus_setup()
        d = self.info(ring)["d"]
        n, kappa = f.shape[0], A.shape[0]
        tau, fc, cm = np.zeros(n, dtype=np.uint64), np.zeros((3, kappa, d), dtype=np.uint64), np.zeros((k, kappa, d, d), dtype=np.uint64)
        self.check(self.lib.lfo_plus_rg_from_f(ring, ptr(np.ascontiguousarray(f)), n, ptr(np.ascontiguousarray(A)), kappa, b, k, l, ptr(tau), ptr(fc), ptr(cm)))
        return tau, fc, cm

    def plus_range_check(self, ring, nvars, fs, A, b, k, l, M=(), seed=None):
        """fs: L x n x d witnesses; A: kappa x n x d.  Returns the Dcom image (from_f of every instance + range_check)."""
        self._plus_setup()
        fs, A = np.ascontiguousarray(fs), np.ascontiguousarray(A)
        ma = make_csr_array(list(M))
        sd, sp, sn = self._seed(seed)
        return self._plus_call(lambda out, cap: self.lib.lfo_plus_range_check(ring, nvars, fs.shape[0], ptr(fs), fs.shape[1], ptr(A), A.shape[0], b, k, l, ma, len(M), sp, sn, ptr(out), cap))

    def plus_range_check_verify(self, ring, words, seed=None):
        self._plus_setup()
        words = np.ascontiguousarray(words, dtype=np.uint64)
        sd, sp, sn = self._seed(seed)
        rc = self.lib.lfo_plus_range_check_verify(ring, ptr(words), words.size, sp, sn)
        if rc < 0:
            raise OracleError(rc, self.err())
        return bool(rc)

    @staticmethod
    def plus_comx_words(nvars, L, kappa, n_M, d=16):
        return L * kappa * d + 2 * nvars + L * (1 + n_M) * 2 * d

    def plus_cm_prove(self, ring, nvars, fs, A, b, k, l, M=(), seed=None, want_g=True):
        """Cm::prove on from_f instances: returns (CmProof image, ComX image, g[L, n, d] or None)."""
        self._plus_setup()
        fs, A = np.ascontiguousarray(fs), np.ascontiguousarray(A)
        ma = make_csr_array(list(M))
        sd, sp, sn = self._seed(seed)
        comx = np.zeros(self.plus_comx_words(nvars, fs.shape[0], A.shape[0], len(M)), dtype=np.uint64)
        g = np.zeros_like(fs) if want_g else None
        proof = self._plus_call(lambda out, cap: self.lib.lfo_plus_cm_prove(ring, nvars, fs.shape[0], ptr(fs), fs.shape[1], ptr(A), A.shape[0], b, k, l, ma, len(M), sp, sn,
                                                                             ptr(out), cap, ptr(comx), comx.size, ptr(g)))
        return proof, comx, g

    def plus_cm_verify(self, ring, words, M=(), seed=None, nvars=None, L=1, kappa=1):
        """CmProof::verify: (accepted, ComX image or None)"""
        self._plus_setup()
        words = np.ascontiguousarray(words, dtype=np.uint64)
        ma = make_csr_array(list(M))
        sd, sp, sn = self._seed(seed)
        comx = np.zeros(self.plus_comx_words(nvars, L, kappa, len(M)), dtype=np.uint64) if nvars else None
        rc = self.lib.lfo_plus_cm_verify(ring, ptr(words), words.size, ma, len(M), sp, sn, ptr(comx), comx.size if comx is not None else 0)
        if rc < 0:
            raise OracleError(rc, self.err())
        return bool(rc), (comx if rc else None)

    def plus_mlin(self, ring, fs, A, b, k, l, M=(), seed=None):
        """Mlin::mlin: (CmProof image, dict(cm_g[kappa, d], ro[nvars, 2], vo[1 + n_M, 2, d]), g[n, d])"""
        self._plus_setup()
        fs, A = np.ascontiguousarray(fs), np.ascontiguousarray(A)
        Lc, n, d = fs.shape; kappa, nE, nv = A.shape[0], 1 + len(M), int(n - 1).bit_length()
        ma = make_csr_array(list(M)); sd, sp, sn = self._seed(seed)
        x = np.zeros(kappa * d + 2 * nv + nE * 2 * d, dtype=np.uint64); g = np.zeros((n, d), dtype=np.uint64)
        proof = self._plus_call(lambda out, cap: self.lib.lfo_plus_mlin(ring, Lc, ptr(fs), n, ptr(A), kappa, b, k, l, ma, len(M), sp, sn, ptr(out), cap, ptr(x), ptr(g)))
        return proof, dict(cm_g=x[: kappa * d].reshape(kappa, d).copy(), ro=x[kappa * d: kappa * d + 2 * nv].reshape(nv, 2).copy(), vo=x[kappa * d + 2 * nv:].reshape(nE, 2, d).copy()), g

    def plus_decompose(self, ring, f, r_pairs, A, B, M=()):
        """Decomp::decompose: (proof words = C0 | C1 | v0 | v1, F[2, n, d])"""
        self._plus_setup()
        f, A, r_pairs = np.ascontiguousarray(f), np.ascontiguousarray(A), np.ascontiguousarray(r_pairs, dtype=np.uint64)
        n, d = f.shape; kappa, nE = A.shape[0], 1 + len(M)
        ma = make_csr_array(list(M))
        proof = np.zeros(2 * kappa * d + 2 * nE * 2 * d, dtype=np.uint64); F = np.zeros((2, n, d), dtype=np.uint64)
        self.check(self.lib.lfo_plus_decompose(ring, ptr(f), n, ptr(r_pairs), ma, len(M), ptr(A), kappa, B, ptr(proof), ptr(F)))
        return proof, F

    def plus_decompose_verify(self, ring, proof, kappa, n_M, cm_f, v, B):
        self._plus_setup()
        rc = self.lib.lfo_plus_decompose_verify(ring, ptr(np.ascontiguousarray(proof)), kappa, n_M, ptr(np.ascontiguousarray(cm_f)), ptr(np.ascontiguousarray(v)), B)
        if rc < 0:
            raise OracleError(rc, self.err())
        return bool(rc)

    # ---- stateful LatticeFold+ transcript and the entry points that take it (the flow of plus.rs)
    def plus_transcript(self, ring, seed=None):
        self._plus_setup()
        sd, sp, sn = self._seed(seed)
        return self.lib.lfo_plus_tr_new(ring, sp, sn)

    def plus_transcript_free(self, tr):
        self.lib.lfo_plus_tr_free(tr)

    def plus_transcript_challenge(self, tr):
        return int(self.lib.lfo_plus_tr_challenge(tr))

    def plus_r1cs_linearize(self, ring, abc, f, tr):
        """ComR1CS::linearize: (LinB dict(f, r, v), proof image)"""
        self._plus_setup()
        f = np.ascontiguousarray(f); ma = make_csr_array(list(abc)); d = f.shape[1]
        img = self._plus_call(lambda out, cap: self.lib.lfo_plus_r1cs_linearize(ring, ma, ptr(f), f.shape[0], tr, ptr(out), cap))
        nv = int(img[0]); ro, v4 = img[1: 1 + nv], img[-4 * d:].reshape(4, d)
        return dict(f=f, r=np.stack([ro, ro], axis=1), v=np.stack([v4, v4], axis=1)), img

    def plus_r1cs_linearize_verify(self, ring, words, tr):
        self._plus_setup()
        words = np.ascontiguousarray(words, dtype=np.uint64)
        rc = self.lib.lfo_plus_r1cs_linearize_verify(ring, ptr(words), words.size, tr)
        if rc < 0:
            raise OracleError(rc, self.err())
        return bool(rc)

    def plus_mlin_t(self, ring, fs, A, b, k, l, M, tr):
        self._plus_setup()
        fs, A = np.ascontiguousarray(fs), np.ascontiguousarray(A)
        Lc, n, d = fs.shape; kappa, nE, nv = A.shape[0], 1 + len(M), int(n - 1).bit_length()
        ma = make_csr_array(list(M))
        x = np.zeros(kappa * d + 2 * nv + nE * 2 * d, dtype=np.uint64); g = np.zeros((n, d), dtype=np.uint64)
        proof = self._plus_call(lambda out, cap: self.lib.lfo_plus_mlin_t(ring, Lc, ptr(fs), n, ptr(A), kappa, b, k, l, ma, len(M), tr, ptr(out), cap, ptr(x), ptr(g)))
        return proof, dict(cm_g=x[: kappa * d].reshape(kappa, d).copy(), ro=x[kappa * d: kappa * d + 2 * nv].reshape(nv, 2).copy(), vo=x[kappa * d + 2 * nv:].reshape(nE, 2, d).copy()), g

    def plus_cm_verify_t(self, ring, words, n_M, tr):
        self._plus_setup()
        words = np.ascontiguousarray(words, dtype=np.uint64)
        rc = self.lib.lfo_plus_cm_verify_t(ring, ptr(words), words.size, n_M, tr)
        if rc < 0:
            raise OracleError(rc, self.err())
        return bool(rc)

    def plus_mat_vec(self, ring, A, x):
        self._plus_setup()
        A, x = np.ascontiguousarray(A), np.ascontiguousarray(x)
        out = np.zeros((A.shape[0], A.shape[2]), dtype=np.uint64)
        self.check(self.lib.lfo_plus_mat_vec(ring, ptr(A), A.shape[0], A.shape[1], ptr(x), ptr(out)))
        return out

    def plus_tensor(self, ring, r):
        self._plus_setup()
        r = np.ascontiguousarray(r, dtype=np.uint64)
        out = np.zeros(1 << r.size, dtype=np.uint64)
        self.check(self.lib.lfo_plus_tensor(ring, ptr(r), r.size, ptr(out)))
        return out

    def plus_ring_mul(self, ring, a, b):
        self._plus_setup()
        out = np.zeros_like(a)
        self.check(self.lib.lfo_plus_ring_mul(ring, ptr(np.ascontiguousarray(a)), ptr(np.ascontiguousarray(b)), ptr(out)))
        return out


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"oracle error {code}: {msg}")
        self.code = code


class OracleTranscript:
    def __init__(self, orc, ring, h=None):
        self.o, self.ring = orc, ring
        self.h = h if h is not None else orc.lib.lfo_tr_new(ring)
        self.d, self.tau = orc.info(ring)["d"], orc.info(ring)["tau"]

    def clone(self):
        return OracleTranscript(self.o, self.ring, self.o.lib.lfo_tr_clone(self.h))

    def __del__(self):
        try:
            self.o.lib.lfo_tr_free(self.h)
        except Exception:
            pass

    def absorb(self, els):
        els = np.ascontiguousarray(els, dtype=np.uint64); self.o.lib.lfo_tr_absorb(self.h, ptr(els), els.size // self.d)

    def absorb_base(self, limbs):
        limbs = np.ascontiguousarray(limbs, dtype=np.uint64); self.o.lib.lfo_tr_absorb_base(self.h, ptr(limbs), limbs.size)

    def absorb_tag(self, tag):
        self.o.lib.lfo_tr_absorb_tag(self.h, tag.encode())

    def squeeze_base(self, n):
        o = np.empty(n, dtype=np.uint64); self.o.lib.lfo_tr_squeeze_base(self.h, ptr(o), n); return o

    def get_challenge(self):
        o = np.empty(self.tau, dtype=np.uint64); self.o.lib.lfo_tr_get_challenge(self.h, ptr(o)); return o

    def get_short_challenge(self):
        o = np.empty(self.d, dtype=np.uint64); self.o.lib.lfo_tr_get_short_challenge(self.h, ptr(o)); return o

    def state(self):
        o = np.empty(24, dtype=np.uint64); self.o.lib.lfo_tr_state(self.h, ptr(o)); return o
