"""Host-side mirror of the LatticeFold+ operators built on the same kernels (crates/latticefold-plus), over the C ABI.

Names follow the reference: `In.set_check` / `Out.verify` (setchk.rs), `RgInstance.from_f`, `Rg.range_check`, `Dcom.verify`
(rgchk.rs), `tensor` (utils.rs), `PoseidonTranscript` (transcript.rs).  The ring is the Frog ring in coefficient form
(`frog_ring::RqPoly`): elements are arrays of 16 canonical uint64 coefficients.  Proof objects are the flat uint64 images
documented in include/lf_b200.h.  No CPU fallback: the provers need the CUDA library and a device; the verifiers are host code."""
import ctypes as C

import numpy as np

from . import synth
from .api import Csr, LfError, Transcript, lib, ptr, u64p, vp

RING = synth.RING_FROG
D = 16
SYMBOLS = """lf_transcript_get_challenge_base lf_plus_set_check lf_plus_set_check_verify lf_plus_mat_create lf_plus_mat_free
lf_plus_rg_from_f lf_plus_rg_read lf_plus_rg_free lf_plus_range_check lf_plus_range_check_verify lf_plus_tensor
lf_plus_comx_words lf_plus_cm_prove lf_plus_cm_verify lf_plus_mlin lf_plus_decompose lf_plus_decompose_verify
lf_plus_r1cs_linearize lf_plus_r1cs_linearize_verify lf_plus_csr_pin lf_plus_csr_unpin
lf_plus_vec_upload lf_plus_vec_download lf_plus_vec_len lf_plus_vec_free lf_plus_r1cs_linearize_v lf_plus_mlin_v lf_plus_decompose_v""".split()


class PlusSet(C.Structure):      # lf_plus_set
    _fields_ = [("kind", C.c_int32), ("pad", C.c_int32), ("m", Csr), ("v", u64p), ("n", C.c_uint64)]


_ready = False


def _L():
    global _ready
    L = lib()
    if not _ready:
        L.lf_transcript_get_challenge_base.argtypes = [vp, u64p]
        L.lf_plus_set_check.argtypes = [vp, vp, C.c_int32, C.POINTER(PlusSet), C.c_int32, C.POINTER(Csr), C.c_int32, u64p, C.c_uint64, u64p]
        L.lf_plus_set_check_verify.argtypes = [vp, u64p, C.c_uint64]
        L.lf_plus_mat_create.argtypes = [vp, C.c_uint64, C.c_uint64, u64p, C.POINTER(vp)]
        L.lf_plus_mat_free.argtypes = [vp, vp]
        L.lf_plus_rg_from_f.argtypes = [vp, vp, u64p, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, C.POINTER(vp)]
        L.lf_plus_rg_read.argtypes = [vp, u64p, u64p, u64p]
        L.lf_plus_rg_free.argtypes = [vp, vp]
        L.lf_plus_range_check.argtypes = [vp, vp, C.c_int32, C.POINTER(vp), C.c_int32, C.POINTER(Csr), C.c_int32, u64p, C.c_uint64, u64p]
        L.lf_plus_range_check_verify.argtypes = [vp, u64p, C.c_uint64]
        L.lf_plus_tensor.argtypes = [u64p, C.c_int32, u64p]
        L.lf_plus_comx_words.restype = C.c_uint64
        L.lf_plus_comx_words.argtypes = [C.c_int32, C.c_int32, C.c_uint64, C.c_int32]
        L.lf_plus_cm_prove.argtypes = [vp, vp, C.c_int32, C.POINTER(vp), C.c_int32, C.POINTER(Csr), C.c_int32, u64p, C.c_uint64, u64p, u64p, u64p]
        L.lf_plus_cm_verify.argtypes = [vp, u64p, C.c_uint64, C.c_int32, u64p]
        L.lf_plus_mlin.argtypes = [vp, vp, vp, u64p, C.c_int32, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, C.POINTER(Csr), C.c_int32, u64p, C.c_uint64, u64p, u64p, u64p]
        L.lf_plus_decompose.argtypes = [vp, vp, u64p, C.c_uint64, u64p, C.POINTER(Csr), C.c_int32, C.c_uint64, u64p, u64p]
        L.lf_plus_decompose_verify.argtypes = [u64p, C.c_uint64, C.c_int32, u64p, u64p, C.c_uint64]
        L.lf_plus_r1cs_linearize.argtypes = [vp, vp, C.POINTER(Csr), u64p, C.c_uint64, u64p, C.c_uint64, u64p]
        L.lf_plus_r1cs_linearize_verify.argtypes = [vp, u64p, C.c_uint64]
        L.lf_plus_csr_pin.argtypes = L.lf_plus_csr_unpin.argtypes = [vp, C.POINTER(Csr)]
        L.lf_plus_vec_upload.argtypes = [vp, u64p, C.c_uint64, C.POINTER(vp)]
        L.lf_plus_vec_download.argtypes = [vp, vp, u64p]
        L.lf_plus_vec_len.restype = C.c_uint64
        L.lf_plus_vec_len.argtypes = [vp]
        L.lf_plus_vec_free.argtypes = [vp, vp]
        L.lf_plus_r1cs_linearize_v.argtypes = [vp, vp, C.POINTER(Csr), vp, u64p, C.c_uint64, u64p]
        L.lf_plus_mlin_v.argtypes = [vp, vp, vp, C.POINTER(vp), C.c_int32, C.c_uint64, C.c_int32, C.c_int32, C.POINTER(Csr), C.c_int32, u64p, C.c_uint64, u64p, u64p, C.POINTER(vp)]
        L.lf_plus_decompose_v.argtypes = [vp, vp, vp, u64p, C.POINTER(Csr), C.c_int32, C.c_uint64, u64p, C.POINTER(vp), C.POINTER(vp)]
        _ready = True
    return L


class PoseidonTranscript(Transcript):
    """latticefold-plus/src/transcript.rs: the same sponge; a challenge is one base-field element."""

    def __init__(self, handle=None):
        super().__init__(RING, handle)
        _L()

    def get_challenge(self):
        o = np.empty(1, dtype=np.uint64); self.L.lf_transcript_get_challenge_base(self.h, ptr(o)); return int(o[0])


def _csr_array(mats):
    arr = (Csr * max(len(mats), 1))()
    for j, M in enumerate(mats):
        arr[j].nrows, arr[j].ncols = M["nrows"], M["ncols"]
        arr[j].row_ptr, arr[j].col, arr[j].val = ptr(M["row_ptr"]), ptr(M["col"]), ptr(M["val"])
    return arr


def _grow(ctx, call):
    cap = 1 << 16
    while True:
        out, n = np.zeros(cap, dtype=np.uint64), C.c_uint64(0)
        rc = call(ptr(out), cap, C.byref(n))
        if rc == 0:
            return out[: n.value].copy()
        if n.value > cap:
            cap = int(n.value); continue
        ctx.check(rc)


class In:
    """setchk.rs:23-27: `sets` is a list of ("matrix", csr dict) / ("vector", n x 16 array)."""

    def __init__(self, ctx, nvars, sets):
        self.ctx, self.nvars, self.sets = ctx, nvars, sets

    def set_check(self, M, transcript):      # setchk.rs:59-262 -> the `Out` image
        L = _L()
        arr = (PlusSet * max(len(self.sets), 1))()
        keep = []
        for i, (kind, x) in enumerate(self.sets):
            if kind == "matrix":
                arr[i].kind = 0
                arr[i].m.nrows, arr[i].m.ncols = x["nrows"], x["ncols"]
                arr[i].m.row_ptr, arr[i].m.col, arr[i].m.val = ptr(x["row_ptr"]), ptr(x["col"]), ptr(x["val"])
            else:
                x = np.ascontiguousarray(x, dtype=np.uint64); keep.append(x)
                arr[i].kind, arr[i].v, arr[i].n = 1, ptr(x), x.shape[0]
        ma = _csr_array(list(M))
        return _grow(self.ctx, lambda o, cap, n: L.lf_plus_set_check(self.ctx.h, transcript.h, self.nvars, arr, len(self.sets), ma, len(M), o, cap, n))


def set_check_verify(words, transcript):      # Out::verify, setchk.rs:264-344: True = Ok(())
    words = np.ascontiguousarray(words, dtype=np.uint64)
    rc = _L().lf_plus_set_check_verify(transcript.h, ptr(words), words.size)
    if rc in (0, -10):
        return rc == 0
    raise LfError(rc, "set-check image rejected as malformed")


class Matrix:
    """Matrix<R> kappa x n in coefficient form, device resident (the `A` of RgInstance::from_f)."""

    def __init__(self, ctx, A):
        A = np.ascontiguousarray(A, dtype=np.uint64)
        self.ctx, self.kappa, self.n, self.h = ctx, A.shape[0], A.shape[1], vp()
        ctx.check(_L().lf_plus_mat_create(ctx.h, self.kappa, self.n, ptr(A), C.byref(self.h)))

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                _L().lf_plus_mat_free(self.ctx.h, self.h)
        except Exception:
            pass


class RgInstance:
    """rgchk.rs:41-48; `from_f` (rgchk.rs:259-336) runs the double commitment on the device."""

    def __init__(self, ctx, handle, n, kappa, k):
        self.ctx, self.h, self.n, self.kappa, self.k = ctx, handle, n, kappa, k

    @classmethod
    def from_f(cls, ctx, f, A, b, k, l):
        f = np.ascontiguousarray(f, dtype=np.uint64)
        h = vp()
        ctx.check(_L().lf_plus_rg_from_f(ctx.h, A.h, ptr(f), f.shape[0], b, k, l, C.byref(h)))
        return cls(ctx, h, f.shape[0], A.kappa, k)

    def read(self):
        """(tau[n], fcoms[3, kappa, 16] = cm_f / C_Mf / cm_mtau, comM_f[k, kappa, 16, 16])"""
        tau, fc, cm = np.zeros(self.n, dtype=np.uint64), np.zeros((3, self.kappa, D), dtype=np.uint64), np.zeros((self.k, self.kappa, D, D), dtype=np.uint64)
        self.ctx.check(_L().lf_plus_rg_read(self.h, ptr(tau), ptr(fc), ptr(cm)))
        return tau, fc, cm

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                _L().lf_plus_rg_free(self.ctx.h, self.h)
        except Exception:
            pass


class Rg:
    """rgchk.rs:34-39"""

    def __init__(self, ctx, nvars, instances):
        self.ctx, self.nvars, self.instances = ctx, nvars, instances

    def range_check(self, M, transcript):      # rgchk.rs:75-187 -> the `Dcom` image
        L = _L()
        hs = (vp * len(self.instances))(*[i.h for i in self.instances])
        ma = _csr_array(list(M))
        return _grow(self.ctx, lambda o, cap, n: L.lf_plus_range_check(self.ctx.h, transcript.h, self.nvars, hs, len(self.instances), ma, len(M), o, cap, n))


def range_check_verify(words, transcript):      # Dcom::verify, rgchk.rs:190-246
    words = np.ascontiguousarray(words, dtype=np.uint64)
    rc = _L().lf_plus_range_check_verify(transcript.h, ptr(words), words.size)
    if rc in (0, -10, -11):
        return rc == 0
    raise LfError(rc, "range-check image rejected as malformed")


class Cm:
    """cm.rs:22-25: the commitment transformation over an `Rg`."""

    def __init__(self, rg):
        self.rg = rg

    def prove(self, M, transcript, want_g=True, g_out=None):      # cm.rs:57-203 -> (CmProof image, ComX image, g[L, n, 16] or None)
        L_, rg = _L(), self.rg
        hs = (vp * len(rg.instances))(*[i.h for i in rg.instances])
        ma = _csr_array(list(M))
        i0 = rg.instances[0]
        comx = np.zeros(int(L_.lf_plus_comx_words(rg.nvars, len(rg.instances), i0.kappa, len(M))), dtype=np.uint64)
        g = g_out if g_out is not None else (np.zeros((len(rg.instances), i0.n, D), dtype=np.uint64) if want_g else None)
        proof = _grow(rg.ctx, lambda o, cap, n: L_.lf_plus_cm_prove(rg.ctx.h, transcript.h, rg.nvars, hs, len(rg.instances), ma, len(M), o, cap, n, ptr(comx), ptr(g)))
        return proof, comx, g


def cm_verify(words, n_M, transcript, nvars=None, L=1, kappa=1):      # CmProof::verify, cm.rs:349-535 -> (accepted, ComX image)
    words = np.ascontiguousarray(words, dtype=np.uint64)
    comx = np.zeros(int(_L().lf_plus_comx_words(nvars, L, kappa, n_M)), dtype=np.uint64) if nvars else None
    rc = _L().lf_plus_cm_verify(transcript.h, ptr(words), words.size, n_M, ptr(comx))
    if rc in (0, -10, -11):
        return rc == 0, (comx if rc == 0 else None)
    raise LfError(rc, "cm proof image rejected as malformed")


class Vec:
    """Vec<R> resident on the device (n x 16 coefficients): the witness of a LinB between the sub-protocols of PlusProver::prove."""

    def __init__(self, ctx, host=None, handle=None):
        self.ctx, self.h = ctx, handle
        if handle is None:
            host = np.ascontiguousarray(host, dtype=np.uint64)
            self.h = vp()
            ctx.check(_L().lf_plus_vec_upload(ctx.h, ptr(host), host.shape[0], C.byref(self.h)))

    def __len__(self):
        return int(_L().lf_plus_vec_len(self.h))

    def download(self):
        out = np.zeros((len(self), D), dtype=np.uint64)
        self.ctx.check(_L().lf_plus_vec_download(self.ctx.h, self.h, ptr(out)))
        return out

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                _L().lf_plus_vec_free(self.ctx.h, self.h)
        except Exception:
            pass


class ComR1CS:
    """r1cs.rs:19-35: R1CS matrices (A, B, C as csr dicts over the committed, decomposed witness) and the witness f (n x 16)."""

    def __init__(self, ctx, abc, f):
        self.ctx, self.abc, self.f = ctx, list(abc), np.ascontiguousarray(f, dtype=np.uint64)
        self._vec = None

    def vec(self):      # the witness on the device, uploaded once
        if self._vec is None:
            self._vec = Vec(self.ctx, self.f)
        return self._vec

    def linearize_resident(self, transcript):      # the same proof from the resident witness -> (Vec, proof image)
        ma = _csr_array(self.abc)
        v = self.vec()
        img = _grow(self.ctx, lambda o, cap, n: _L().lf_plus_r1cs_linearize_v(self.ctx.h, transcript.h, ma, v.h, o, cap, n))
        return v, img

    def matrices(self):
        return self.abc

    def linearize(self, transcript):      # r1cs.rs:72-134 -> (LinB dict(f, r, v), proof image)
        ma = _csr_array(self.abc)
        img = _grow(self.ctx, lambda o, cap, n: _L().lf_plus_r1cs_linearize(self.ctx.h, transcript.h, ma, ptr(self.f), self.f.shape[0], o, cap, n))
        nv = int(img[0])
        ro, v4 = img[1: 1 + nv], img[-4 * D:].reshape(4, D)
        return dict(f=self.f, r=np.stack([ro, ro], axis=1), v=np.stack([v4, v4], axis=1)), img


def r1cs_linearize_verify(words, transcript):      # ComR1CSProof::verify, r1cs.rs:136-162
    words = np.ascontiguousarray(words, dtype=np.uint64)
    rc = _L().lf_plus_r1cs_linearize_verify(transcript.h, ptr(words), words.size)
    if rc in (0, -10):
        return rc == 0
    raise LfError(rc, "linearization image rejected as malformed")


class Mlin:
    """mlin.rs:15-19: L witnesses `fs` (L x n x 16) folded by Cm::prove; `mlin` returns (CmProof image, LinB2X dict, g[n, 16])."""

    def __init__(self, ctx, fs, b, k, l):
        self.ctx, self.fs, self.b, self.k, self.l = ctx, np.ascontiguousarray(fs, dtype=np.uint64), b, k, l

    def mlin(self, A, M, transcript, want_g=True):      # mlin.rs:41-106
        L_ = _L()
        Lc, n, d = self.fs.shape; nE, nv = 1 + len(M), int(n - 1).bit_length()
        ma = _csr_array(list(M))
        x = np.zeros(A.kappa * d + 2 * nv + nE * 2 * d, dtype=np.uint64)
        g = np.zeros((n, d), dtype=np.uint64) if want_g else None
        proof = _grow(self.ctx, lambda o, cap, ln: L_.lf_plus_mlin(self.ctx.h, transcript.h, A.h, ptr(self.fs), Lc, n, self.b, self.k, self.l, ma, len(M), o, cap, ln, ptr(x), ptr(g)))
        kd = A.kappa * d
        return proof, dict(cm_g=x[:kd].reshape(A.kappa, d).copy(), ro=x[kd: kd + 2 * nv].reshape(nv, 2).copy(), vo=x[kd + 2 * nv:].reshape(nE, 2, d).copy()), g


def decompose(ctx, A, f, r_pairs, M, B, want_F=True):      # Decomp::decompose, decomp.rs:32-99 -> (DecompProof words, F[2, n, 16])
    f, r_pairs = np.ascontiguousarray(f, dtype=np.uint64), np.ascontiguousarray(r_pairs, dtype=np.uint64)
    n, d = f.shape
    ma = _csr_array(list(M))
    proof = np.zeros(2 * A.kappa * d + 2 * (1 + len(M)) * 2 * d, dtype=np.uint64)
    F = np.zeros((2, n, d), dtype=np.uint64) if want_F else None
    ctx.check(_L().lf_plus_decompose(ctx.h, A.h, ptr(f), n, ptr(r_pairs), ma, len(M), B, ptr(proof), ptr(F)))
    return proof, F


def decompose_verify(proof, kappa, n_M, cm_f, v, B):      # DecompProof::verify, decomp.rs:102-126
    rc = _L().lf_plus_decompose_verify(ptr(np.ascontiguousarray(proof, dtype=np.uint64)), kappa, n_M, ptr(np.ascontiguousarray(cm_f, dtype=np.uint64)), ptr(np.ascontiguousarray(v, dtype=np.uint64)), B)
    if rc in (0, -11):
        return rc == 0
    raise LfError(rc, "decomposition proof rejected as malformed")


def estimate_bound(sop, L, d, k):      # utils.rs:105-115
    a = sop * L
    c = d // 2 + d * k + 1
    return int(np.ceil((a + np.sqrt(float(a * a + 4 * a * c))) / 2.0))


def sparse_gadget_decompose(M, b, k, p=15912092521325583641):
    """SparseMatrix::gadget_decompose(b, k): column j becomes the k columns j k + t with entries M[i][j] b^t, so that the result times
    gadget_decompose(z, b, k) equals M z (stark-rings-linalg; the digit order is the one of Vec::gadget_decompose: digit t of element j at j k + t)."""
    nnz = M["col"].size
    col = (M["col"].astype(np.uint64)[:, None] * np.uint64(k) + np.arange(k, dtype=np.uint64)[None, :]).reshape(-1)
    pw = [pow(b, t, p) for t in range(k)]
    val = np.zeros((nnz, k, D), dtype=np.uint64)
    src = M["val"].astype(object)
    for t in range(k):
        val[:, t, :] = ((src * pw[t]) % p).astype(np.uint64)
    row_ptr = (M["row_ptr"].astype(np.uint64) * np.uint64(k)).astype(np.uint64)
    return dict(nrows=M["nrows"], ncols=M["ncols"] * k, row_ptr=np.ascontiguousarray(row_ptr), col=np.ascontiguousarray(col), val=np.ascontiguousarray(val.reshape(nnz * k, D)))


def pad_rows(M, n):      # SparseMatrix::pad_rows: empty rows up to n
    if M["nrows"] >= n:
        return M
    rp = np.concatenate([M["row_ptr"], np.full(n - M["nrows"], M["row_ptr"][-1], dtype=np.uint64)])
    return dict(M, nrows=n, row_ptr=np.ascontiguousarray(rp))


def r1cs_decomposed_square(abc, n, b, k):      # r1cs.rs:171-185: n x m -> n x n with m k = n
    return [pad_rows(sparse_gadget_decompose(M, b, k), n) for M in abc]


def com_r1cs_new(ctx, abc, z, b, k):
    """ComR1CS::new (r1cs.rs:46-58): f = gadget_decompose(z, b, k) on the device; the commitment cm_f = A f is RgInstance::from_f's, not needed here."""
    from .api import FORM_COEFF
    f = ctx.gadget_decompose(ctx.upload(np.ascontiguousarray(z, dtype=np.uint64), FORM_COEFF), b, k).download()
    return ComR1CS(ctx, abc, f)


class PinnedMatrices:
    """lf_plus_csr_pin for a list of csr dicts: resident copies of the static matrices for as long as this object lives."""

    def __init__(self, ctx, mats):
        self.ctx, self.mats, self.arr = ctx, list(mats), _csr_array(list(mats))
        for i in range(len(self.mats)):
            ctx.check(_L().lf_plus_csr_pin(ctx.h, C.byref(self.arr[i])))

    def close(self):
        if self.arr is not None and self.ctx.h:
            for i in range(len(self.mats)):
                _L().lf_plus_csr_unpin(self.ctx.h, C.byref(self.arr[i]))
        self.arr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PlusProver:
    """plus.rs:15-24, 48-118: the accumulating LatticeFold+ prover.  Host orchestration as in the reference; every step it calls runs on the device."""

    def __init__(self, ctx, A, M, b, k, l, B, transcript):      # PlusProver::init
        self.ctx, self.A, self.M, self.b, self.k, self.l, self.B, self.transcript, self.acc = ctx, A, list(M), b, k, l, B, transcript, []
        self.pinned = PinnedMatrices(ctx, self.M)      # the reference's prover owns M as well

    def prove(self, comps):      # plus.rs:80-117 -> PlusProof dict(linb2x, lproof, cmproof, dproof); the witnesses stay on the device between the steps
        L_ = _L(); lproof = []
        for comp in comps:
            v, lp = comp.linearize_resident(self.transcript)
            lproof.append(lp); self.acc.append(v)
        n, nE = len(self.acc[0]), 1 + len(self.M); nv = int(n - 1).bit_length()
        ma = _csr_array(self.M)
        hs = (vp * len(self.acc))(*[v.h for v in self.acc])
        x = np.zeros(self.A.kappa * D + 2 * nv + nE * 2 * D, dtype=np.uint64); gh = vp()
        cmproof = _grow(self.ctx, lambda o, cap, ln: L_.lf_plus_mlin_v(self.ctx.h, self.transcript.h, self.A.h, hs, len(self.acc), self.b, self.k, self.l, ma, len(self.M), o, cap, ln, ptr(x), C.byref(gh)))
        g = Vec(self.ctx, handle=gh); kd = self.A.kappa * D
        linb2x = dict(cm_g=x[:kd].reshape(self.A.kappa, D).copy(), ro=x[kd: kd + 2 * nv].reshape(nv, 2).copy(), vo=x[kd + 2 * nv:].reshape(nE, 2, D).copy())
        dproof = np.zeros(2 * kd + 2 * nE * 2 * D, dtype=np.uint64); f0, f1 = vp(), vp()
        ro = np.ascontiguousarray(linb2x["ro"])
        self.ctx.check(L_.lf_plus_decompose_v(self.ctx.h, self.A.h, g.h, ptr(ro), ma, len(self.M), self.B, ptr(dproof), C.byref(f0), C.byref(f1)))
        self.acc = [Vec(self.ctx, handle=f0), Vec(self.ctx, handle=f1)]      # keep only the accumulated instance
        return dict(linb2x=linb2x, lproof=lproof, cmproof=cmproof, dproof=dproof)

    def acc_download(self):
        return [v.download() for v in self.acc]


class PlusVerifier:
    """plus.rs:26-33, 120-145"""

    def __init__(self, kappa, n_M, B, transcript):
        self.kappa, self.n_M, self.B, self.transcript = kappa, n_M, B, transcript

    def verify(self, proof):
        ok = all(r1cs_linearize_verify(lp, self.transcript) for lp in proof["lproof"])
        ok = ok and cm_verify(proof["cmproof"], self.n_M, self.transcript)[0]
        return ok and decompose_verify(proof["dproof"], self.kappa, self.n_M, proof["linb2x"]["cm_g"], proof["linb2x"]["vo"], self.B)


def tensor(r):      # utils.rs:74-86
    r = np.ascontiguousarray(r, dtype=np.uint64)
    out = np.zeros(1 << r.size, dtype=np.uint64)
    rc = _L().lf_plus_tensor(ptr(r), r.size, ptr(out))
    if rc:
        raise LfError(rc, "tensor")
    return out
