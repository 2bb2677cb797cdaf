"""Host-side mirror of the reference's operator surface over the C ABI (include/lf_b200.h), via ctypes.

Names follow the reference: `AjtaiCommitmentScheme.commit`, `crt`/`icrt`, `gadget_decompose`, `decompose_to_vec`,
`MLSumcheck.prove_as_subprotocol`, `evaluate_mles`, `mat_vec_mul`, `NIFSProver.prove`.  Ring elements cross as numpy
uint64 arrays whose last axis holds D canonical limbs.  There is no CPU fallback: loading fails loudly if the shared
library has not been built, and every compute call fails with LF_ERR_CUDA if there is no device.
"""
import ctypes as C
import os

import numpy as np

from . import synth

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_lib", "liblf_b200.so")
u64p = C.POINTER(C.c_uint64)
i32p = C.POINTER(C.c_int32)
vp = C.c_void_p

COLLECTIVE_FN = C.CFUNCTYPE(C.c_int32, vp, C.c_int32, vp, C.c_size_t)   # lf_collective_fn
LF_COMB_PRODUCTS, LF_COMB_LIN, LF_COMB_FOLD = 0, 1, 2
FORM_NTT, FORM_COEFF = 0, 1

# every symbol include/lf_b200.h declares (tests check that the built library exports all of them)
SYMBOLS = """lf_ring_describe lf_ctx_create lf_ctx_destroy lf_last_error lf_ctx_sync lf_ctx_stream lf_ctx_launches lf_ctx_set_bulk_repr lf_ctx_profile lf_ctx_profile_report lf_ctx_set_shard lf_ctx_collectives lf_nccl_unique_id lf_ctx_set_shard_nccl lf_ctx_p2p_export lf_ctx_p2p_import
lf_vec_upload lf_vec_download lf_vec_len lf_vec_form lf_vec_free lf_crt lf_icrt lf_gadget_decompose lf_gadget_recompose
lf_decompose_to_vec lf_fhat lf_ajtai_create lf_ajtai_free lf_ajtai_kappa lf_ajtai_width lf_commit lf_commit_batch lf_commit_coeff
lf_decompose_and_commit_coeff lf_decompose_and_commit_ntt lf_commit_pieces
lf_sparse_create lf_sparse_free lf_spmv lf_eq_table lf_mle_eval_batch lf_lincomb lf_sumcheck_begin lf_sumcheck_round
lf_sumcheck_finish lf_sumcheck_free lf_transcript_create lf_transcript_clone lf_transcript_free lf_transcript_absorb
lf_transcript_absorb_base lf_transcript_absorb_tag lf_transcript_get_challenge lf_transcript_get_short_challenge
lf_transcript_permutations lf_host_poseidon_backend lf_rot_lin_combination lf_prover_create lf_prover_free lf_proof_words lf_lcccs_words
lf_witness_f_from_w_ccs lf_linearize lf_nifs_prove lf_nifs_verify lf_prover_upload_witness lf_witness_free lf_witness_download_f
lf_proof_wire_bytes lf_proof_serialize lf_proof_deserialize lf_nifs_prove_resident lf_linearization_verify lf_linearize_resident lf_witness_commit lf_prover_last_timings lf_prover_timing_detail
lf_ntt_root lf_ntt_plan_create lf_ntt_plan_free lf_ntt_forward_device lf_ntt_inverse_device lf_ntt_forward_host lf_ntt_inverse_host
lf_ntt_pointwise_mul_device lf_ntt_negacyclic_mul_host""".split()


class LfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"lf_b200 error {code}: {msg}")
        self.code = code


class Csr(C.Structure):
    _fields_ = [("nrows", C.c_uint64), ("ncols", C.c_uint64), ("row_ptr", u64p), ("col", u64p), ("val", u64p)]


class Problem(C.Structure):   # lf_problem
    _fields_ = [("ring", C.c_int32), ("L", C.c_int32), ("K", C.c_int32), ("B_lo", C.c_uint64), ("B_hi", C.c_uint64), ("b", C.c_uint64),
                ("kappa", C.c_uint64), ("n", C.c_uint64), ("A", u64p),
                ("m", C.c_uint64), ("n_ccs", C.c_uint64), ("l", C.c_uint64), ("t", C.c_uint64), ("q", C.c_uint64), ("d", C.c_uint64), ("s", C.c_uint64),
                ("M", C.POINTER(Csr)), ("S_flat", i32p), ("S_len", i32p), ("c", u64p),
                ("acc_r", u64p), ("acc_v", u64p), ("acc_cm", u64p), ("acc_u", u64p), ("acc_x_w", u64p), ("acc_h", u64p),
                ("w_acc_f", u64p), ("cm_i_cm", u64p), ("cm_i_x_ccs", u64p), ("w_i_f", u64p)]


class Comb(C.Structure):      # lf_comb
    _fields_ = [("kind", C.c_int32), ("n_terms", C.c_int32), ("coef_host", u64p), ("idx", i32p), ("idx_len", i32p),
                ("n_mu", C.c_int32), ("b", C.c_int32), ("mu_host", u64p)]


def ptr(a):
    if a is None:
        return None
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags["C_CONTIGUOUS"])
    return a.ctypes.data_as(u64p)


def make_problem(p):
    """dict from synth.make_instance -> (lf_problem, keepalive list)"""
    keep = []
    P = Problem()
    P.ring, P.L, P.K = p["ring"], p["L"], p["K"]
    P.B_lo, P.B_hi, P.b = p["B"] & (2**64 - 1), p["B"] >> 64, p["b"]
    P.kappa, P.n = p["kappa"], p["n"]
    P.A = ptr(p.get("A"))
    ccs = p["ccs"]
    for k in ("m", "n_ccs", "l", "t", "q", "d", "s"):
        setattr(P, k, ccs[k])
    arr = (Csr * ccs["t"])()
    for j, M in enumerate(ccs["M"]):
        arr[j].nrows, arr[j].ncols = M["nrows"], M["ncols"]
        arr[j].row_ptr, arr[j].col, arr[j].val = ptr(M["row_ptr"]), ptr(M["col"]), ptr(M["val"])
    P.M = arr
    S_flat = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=np.int32) for s in ccs["S"]]))
    S_len = np.ascontiguousarray(np.array([len(s) for s in ccs["S"]], dtype=np.int32))
    P.S_flat, P.S_len = S_flat.ctypes.data_as(i32p), S_len.ctypes.data_as(i32p)
    P.c = ptr(ccs["c"])
    acc = p.get("acc")
    if acc is not None:
        P.acc_r, P.acc_v, P.acc_cm, P.acc_u, P.acc_x_w, P.acc_h = (ptr(acc[k]) for k in ("r", "v", "cm", "u", "x_w", "h"))
    P.w_acc_f = ptr(p.get("w_acc_f"))
    P.cm_i_cm, P.cm_i_x_ccs = ptr(p.get("cm_i_cm")), ptr(p["cm_i_x_ccs"])
    P.w_i_f = ptr(p.get("w_i_f"))
    keep += [arr, S_flat, S_len, p]
    return P, keep


_lib = None


def lib():
    """Load liblf_b200.so (built by __graft_entry__.build() / `make -C latticefold_b200/csrc`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`; there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.lf_last_error.restype = C.c_char_p
    L.lf_last_error.argtypes = [vp]
    L.lf_ctx_create.argtypes = [C.c_int32, C.c_int32, C.POINTER(vp)]
    L.lf_ctx_destroy.argtypes = [vp]
    L.lf_ctx_sync.argtypes = [vp]
    L.lf_ctx_stream.restype = vp
    L.lf_ctx_stream.argtypes = [vp]
    L.lf_ctx_launches.restype = C.c_uint64
    L.lf_ctx_launches.argtypes = [vp]
    L.lf_ctx_set_shard.argtypes = [vp, C.c_int32, C.c_int32, COLLECTIVE_FN, vp]
    L.lf_nccl_unique_id.argtypes = [C.POINTER(C.c_uint8)]
    L.lf_ctx_set_shard_nccl.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(C.c_uint8)]
    L.lf_ctx_p2p_export.argtypes = [vp, C.POINTER(C.c_uint8)]
    L.lf_ctx_p2p_import.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(C.c_uint8)]
    L.lf_ctx_collectives.restype = C.c_uint64
    L.lf_ctx_collectives.argtypes = [vp]
    L.lf_ctx_set_bulk_repr.argtypes = [vp, C.c_int32]
    L.lf_ctx_profile.argtypes = [vp, C.c_int32]
    L.lf_ctx_profile_report.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.lf_vec_upload.argtypes = [vp, u64p, C.c_size_t, C.c_int32, C.POINTER(vp)]
    L.lf_vec_download.argtypes = [vp, vp, u64p]
    L.lf_vec_len.restype = C.c_size_t
    L.lf_vec_len.argtypes = [vp]
    L.lf_vec_form.argtypes = [vp]
    L.lf_vec_free.argtypes = [vp, vp]
    L.lf_crt.argtypes = L.lf_icrt.argtypes = [vp, vp, C.POINTER(vp)]
    L.lf_gadget_decompose.argtypes = L.lf_gadget_recompose.argtypes = [vp, vp, C.c_uint64, C.c_int32, C.POINTER(vp)]
    L.lf_decompose_to_vec.argtypes = [vp, vp, C.c_uint64, C.c_int32, C.POINTER(vp)]
    L.lf_fhat.argtypes = [vp, vp, C.POINTER(vp)]
    L.lf_ajtai_create.argtypes = [vp, C.c_size_t, C.c_size_t, u64p, C.POINTER(vp)]
    L.lf_ajtai_free.argtypes = [vp, vp]
    L.lf_ajtai_kappa.restype = L.lf_ajtai_width.restype = C.c_size_t
    L.lf_ajtai_kappa.argtypes = L.lf_ajtai_width.argtypes = [vp]
    L.lf_commit.argtypes = [vp, vp, vp, u64p]
    L.lf_commit_batch.argtypes = [vp, vp, C.POINTER(vp), C.c_int32, u64p]
    L.lf_commit_coeff.argtypes = [vp, vp, vp, u64p]
    L.lf_decompose_and_commit_coeff.argtypes = L.lf_decompose_and_commit_ntt.argtypes = [vp, vp, vp, C.c_uint64, C.c_int32, u64p]
    L.lf_commit_pieces.argtypes = [vp, vp, vp, C.c_uint64, C.c_int32, u64p]
    L.lf_sparse_create.argtypes = [vp, C.c_size_t, C.c_size_t, u64p, u64p, u64p, C.POINTER(vp)]
    L.lf_sparse_free.argtypes = [vp, vp]
    L.lf_spmv.argtypes = [vp, vp, vp, C.POINTER(vp)]
    L.lf_eq_table.argtypes = [vp, u64p, C.c_int32, C.POINTER(vp)]
    L.lf_mle_eval_batch.argtypes = [vp, C.POINTER(vp), C.c_int32, C.c_int32, u64p, C.c_int32, u64p]
    L.lf_lincomb.argtypes = [vp, u64p, C.POINTER(vp), C.c_int32, C.POINTER(vp)]
    L.lf_sumcheck_begin.argtypes = [vp, C.POINTER(vp), C.c_int32, C.c_int32, C.c_int32, C.POINTER(Comb), C.POINTER(vp)]
    L.lf_sumcheck_round.argtypes = [vp, u64p, u64p]
    L.lf_sumcheck_finish.argtypes = [vp, u64p, u64p]
    L.lf_sumcheck_free.argtypes = [vp]
    L.lf_transcript_create.argtypes = [C.c_int32, C.POINTER(vp)]
    L.lf_transcript_clone.argtypes = [vp, C.POINTER(vp)]
    L.lf_transcript_free.argtypes = [vp]
    L.lf_transcript_absorb.argtypes = L.lf_transcript_absorb_base.argtypes = [vp, u64p, C.c_size_t]
    L.lf_transcript_absorb_tag.argtypes = [vp, C.c_char_p]
    L.lf_transcript_get_challenge.argtypes = L.lf_transcript_get_short_challenge.argtypes = [vp, u64p]
    L.lf_transcript_permutations.restype = C.c_uint64
    L.lf_transcript_permutations.argtypes = [vp]
    L.lf_host_poseidon_backend.restype = C.c_char_p
    L.lf_host_poseidon_backend.argtypes = []
    L.lf_rot_lin_combination.argtypes = [C.c_int32, u64p, u64p, C.c_int32, u64p]
    L.lf_prover_create.argtypes = [vp, C.POINTER(Problem), C.POINTER(vp)]
    L.lf_prover_free.argtypes = [vp]
    L.lf_proof_words.restype = L.lf_lcccs_words.restype = C.c_uint64
    L.lf_proof_words.argtypes = L.lf_lcccs_words.argtypes = [C.POINTER(Problem)]
    L.lf_witness_f_from_w_ccs.argtypes = [vp, u64p, C.c_size_t, C.c_uint64, C.c_int32, u64p]
    L.lf_linearize.argtypes = [vp, C.POINTER(Problem), vp, u64p, u64p]
    L.lf_nifs_prove.argtypes = [vp, C.POINTER(Problem), vp, u64p, u64p, u64p]
    L.lf_nifs_verify.argtypes = [C.POINTER(Problem), vp, u64p, u64p]
    L.lf_proof_wire_bytes.restype = C.c_uint64
    L.lf_proof_wire_bytes.argtypes = [C.POINTER(Problem)]
    L.lf_proof_serialize.argtypes = [C.POINTER(Problem), u64p, C.POINTER(C.c_uint8)]
    L.lf_proof_deserialize.argtypes = [C.POINTER(Problem), C.POINTER(C.c_uint8), C.c_uint64, u64p]
    L.lf_linearization_verify.argtypes = [C.POINTER(Problem), vp, u64p, u64p]
    L.lf_linearize_resident.argtypes = [vp, C.POINTER(Problem), vp, vp, u64p, u64p]
    L.lf_witness_commit.argtypes = [vp, vp, u64p]
    L.lf_prover_upload_witness.argtypes = [vp, u64p, C.POINTER(vp)]
    L.lf_witness_free.argtypes = [vp, vp]
    L.lf_witness_download_f.argtypes = [vp, vp, u64p]
    L.lf_nifs_prove_resident.argtypes = [vp, C.POINTER(Problem), vp, vp, vp, u64p, u64p, C.POINTER(vp)]
    L.lf_prover_last_timings.argtypes = [vp, C.POINTER(C.c_double)]
    L.lf_prover_timing_detail.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.lf_ntt_root.argtypes = [C.c_int32, C.c_int32, u64p]
    L.lf_ntt_plan_create.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(vp)]
    L.lf_ntt_plan_free.argtypes = [vp, vp]
    for f in ("lf_ntt_forward_device", "lf_ntt_inverse_device", "lf_ntt_forward_host", "lf_ntt_inverse_host"):
        getattr(L, f).argtypes = [vp, vp, vp, vp, C.c_size_t]
    L.lf_ntt_pointwise_mul_device.argtypes = L.lf_ntt_negacyclic_mul_host.argtypes = [vp, vp, vp, vp, vp, C.c_size_t]
    _lib = L
    return L


class Transcript:
    """PoseidonTranscript (crates/latticefold/src/transcript/poseidon.rs:18-75), host side."""

    def __init__(self, ring, handle=None):
        self.L, self.ring = lib(), ring
        R = synth.RINGS[ring]
        self.d, self.tau = R["d"], R["tau"]
        if handle is None:
            h = vp()
            rc = self.L.lf_transcript_create(ring, C.byref(h))
            if rc:
                raise LfError(rc, self.L.lf_last_error(None).decode())
            handle = h
        self.h = handle

    def clone(self):
        h = vp(); self.L.lf_transcript_clone(self.h, C.byref(h)); return Transcript(self.ring, h)

    def __del__(self):
        try:
            self.L.lf_transcript_free(self.h)
        except Exception:
            pass

    def absorb(self, els):
        els = np.ascontiguousarray(els, dtype=np.uint64); self.L.lf_transcript_absorb(self.h, ptr(els), els.size // self.d)

    def absorb_base(self, limbs):
        limbs = np.ascontiguousarray(limbs, dtype=np.uint64); self.L.lf_transcript_absorb_base(self.h, ptr(limbs), limbs.size)

    def absorb_tag(self, tag):
        self.L.lf_transcript_absorb_tag(self.h, tag.encode())

    def get_challenge(self):
        o = np.empty(self.tau, dtype=np.uint64); self.L.lf_transcript_get_challenge(self.h, ptr(o)); return o

    def get_short_challenge(self):
        o = np.empty(self.d, dtype=np.uint64); self.L.lf_transcript_get_short_challenge(self.h, ptr(o)); return o

    def permutations(self):
        return int(self.L.lf_transcript_permutations(self.h))

    def backend(self):
        """dense-layer implementation of the host Poseidon on this machine ("avx512-ifma" / "scalar")"""
        return self.L.lf_host_poseidon_backend().decode()


class DeviceVec:
    """Vec<R> resident in HBM."""

    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle

    def __len__(self):
        return self.ctx.L.lf_vec_len(self.h)

    def form(self):
        return self.ctx.L.lf_vec_form(self.h)

    def download(self):
        o = np.empty((len(self), self.ctx.d), dtype=np.uint64)
        self.ctx.check(self.ctx.L.lf_vec_download(self.ctx.h, self.h, ptr(o)))
        return o

    def release(self):
        h, self.h = self.h, None
        return h

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.L.lf_vec_free(self.ctx.h, self.h)
        except Exception:
            pass


class Context:
    """One (GPU, ring) pair; owns a CUDA stream."""

    def __init__(self, ring=synth.RING_GOLDILOCKS, device=0):
        self.L = lib()
        self.ring = ring
        R = synth.RINGS[ring]
        self.d, self.tau, self.S, self.p = R["d"], R["tau"], R["S"], R["p"]
        self.device, self.rank, self.world = device, 0, 1
        h = vp()
        rc = self.L.lf_ctx_create(ring, device, C.byref(h))
        if rc:
            raise LfError(rc, self.L.lf_last_error(None).decode())
        self.h = h

    def close(self):
        if self.h:
            self.L.lf_ctx_destroy(self.h); self.h = None

    def check(self, rc):
        if rc:
            raise LfError(rc, self.L.lf_last_error(self.h).decode())

    def check_global(self, rc):
        if rc:
            raise LfError(rc, self.L.lf_last_error(None).decode())

    def sync(self):
        self.check(self.L.lf_ctx_sync(self.h))

    def launches(self):
        return int(self.L.lf_ctx_launches(self.h))

    def stream(self):
        return self.L.lf_ctx_stream(self.h)

    def set_shard(self, rank, world, group=None):
        """Shard the witness-column / hypercube axis over `world` ranks (one Context per rank).  Collectives go through
        torch.distributed on `group` (NCCL between GPUs; gloo works too, staged through host memory)."""
        from . import parallel
        import torch.distributed as dist
        if world > 1 and dist.get_backend(group) == "nccl":
            # the library's own communicator: collectives are enqueued on the context's stream (no host round trip)
            import torch
            ident = (C.c_uint8 * 128)()
            if rank == 0:
                self.check_global(self.L.lf_nccl_unique_id(ident))
            t = torch.tensor(list(ident), dtype=torch.uint8, device=f"cuda:{self.device}")
            dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            ident = (C.c_uint8 * 128)(*t.cpu().tolist())
            self.check(self.L.lf_ctx_set_shard_nccl(self.h, rank, world, ident))
            self.rank, self.world = rank, world
            if os.environ.get("LF_P2P", "1") != "0":
                self._enable_p2p(rank, world, group, torch, dist)
        else:
            self._coll = parallel.make_collective(self, group)      # keep the ctypes callback alive
            self.check(self.L.lf_ctx_set_shard(self.h, rank, world, self._coll, None))
        self.rank, self.world = rank, world

    def _enable_p2p(self, rank, world, group, torch, dist):
        """Map every rank's mailbox region (CUDA IPC) so that the small all-reduces run as one kernel with NVLink peer stores.
        All ranks must agree: if any rank cannot export or import, everyone stays on the NCCL path."""
        hbuf = (C.c_uint8 * 64)()
        ok = self.L.lf_ctx_p2p_export(self.h, hbuf) == 0
        mine = torch.tensor(list(hbuf) + [1 if ok else 0], dtype=torch.uint8, device=f"cuda:{self.device}")
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine, group=group)
        allh = torch.stack(allh).cpu()
        self.p2p = False
        if bool(allh[:, 64].all()):
            flat = (C.c_uint8 * (64 * world))(*allh[:, :64].reshape(-1).tolist())
            ok = self.L.lf_ctx_p2p_import(self.h, rank, world, flat) == 0
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=f"cuda:{self.device}")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 1:
                self.p2p = True
            else:
                self.L.lf_ctx_p2p_import(self.h, rank, world, None)

    def set_bulk_repr(self, montgomery):
        """witness-sized host vectors are arkworks Montgomery limbs (a * 2^64 mod p) instead of canonical ones"""
        self.check(self.L.lf_ctx_set_bulk_repr(self.h, 1 if montgomery else 0))

    def collectives(self):
        return int(self.L.lf_ctx_collectives(self.h))

    def profile(self, enable):
        self.check(self.L.lf_ctx_profile(self.h, 1 if enable else 0))

    def profile_report(self):
        """{kernel: (launches, total_ms)} from the event pairs recorded since profile(True)."""
        buf = C.create_string_buffer(1 << 16)
        self.check(self.L.lf_ctx_profile_report(self.h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.split()
            out[name] = (int(cnt), float(ms))
        return out

    # ---- vectors
    def upload(self, a, form=FORM_NTT):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        h = vp(); self.check(self.L.lf_vec_upload(self.h, ptr(a), a.size // self.d, form, C.byref(h)))
        return DeviceVec(self, h)

    def _unary(self, fn, v, *args):
        h = vp(); self.check(fn(self.h, v.h, *args, C.byref(h))); return DeviceVec(self, h)

    def crt(self, v):           # CRT::elementwise_crt
        return self._unary(self.L.lf_crt, v)

    def icrt(self, v):          # ICRT::elementwise_icrt
        return self._unary(self.L.lf_icrt, v)

    def gadget_decompose(self, v, B, L):
        return self._unary(self.L.lf_gadget_decompose, v, B, L)

    def gadget_recompose(self, v, B, L):
        return self._unary(self.L.lf_gadget_recompose, v, B, L)

    def decompose_to_vec(self, v, b, K):
        arr = (vp * K)(); self.check(self.L.lf_decompose_to_vec(self.h, v.h, b, K, arr)); return [DeviceVec(self, vp(x)) for x in arr]

    def fhat(self, v):          # Witness::get_fhat
        arr = (vp * self.tau)(); self.check(self.L.lf_fhat(self.h, v.h, arr)); return [DeviceVec(self, vp(x)) for x in arr]

    def eq_table(self, r):      # build_eq_x_r
        r = np.ascontiguousarray(r, dtype=np.uint64); h = vp()
        self.check(self.L.lf_eq_table(self.h, ptr(r), r.size // self.d, C.byref(h))); return DeviceVec(self, h)

    def evaluate_mles(self, mles, num_vars, point):   # utils/mle_helpers.rs:65-88
        point = np.ascontiguousarray(point, dtype=np.uint64); arr = (vp * len(mles))(*[m.h for m in mles])
        o = np.empty((len(mles), self.d), dtype=np.uint64)
        self.check(self.L.lf_mle_eval_batch(self.h, arr, len(mles), num_vars, ptr(point), point.size // self.d, ptr(o))); return o

    def lincomb(self, coeffs, vecs):
        coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64); arr = (vp * len(vecs))(*[m.h for m in vecs]); h = vp()
        self.check(self.L.lf_lincomb(self.h, ptr(coeffs), arr, len(vecs), C.byref(h))); return DeviceVec(self, h)

    def witness_f_from_w_ccs(self, ring, w_ccs, B, L):   # Witness::from_w_ccs -> f  (ops interface of synth.make_instance)
        w = np.ascontiguousarray(w_ccs, dtype=np.uint64); W = w.size // self.d; o = np.empty((W * L, self.d), dtype=np.uint64)
        self.check(self.L.lf_witness_f_from_w_ccs(self.h, ptr(w), W, B, L, ptr(o))); return o

    def ntt_mul(self, ring, a, b):                        # ops interface: slot-wise product (test-sized; host integers, nu from the library)
        class Info(C.Structure):
            _fields_ = [("p", C.c_uint64), ("d", C.c_int32), ("n_slots", C.c_int32), ("tau", C.c_int32), ("nu", C.c_uint64)]
        info = Info(); self.check_global(self.L.lf_ring_describe(ring, C.byref(info)))
        return synth.sf_mul(ring, np.asarray(a, dtype=np.uint64), np.asarray(b, dtype=np.uint64), int(info.nu))

    def commit(self, ring, A, f):                         # ops interface: one-shot commit from host arrays
        sch = AjtaiCommitmentScheme(self, A); return sch.commit(self.upload(f))

    def linearize(self, prob):                            # ops interface: accumulator = linearization of the instance
        pr = NIFSProver(self, prob); lc, _ = pr.linearize(prob, Transcript(self.ring)); pr.close()
        return synth.split_lcccs(self.ring, prob, lc)


FIELD_GOLDILOCKS, FIELD_BABYBEAR = 0, 1
NTT_DTYPE = {FIELD_GOLDILOCKS: np.uint64, FIELD_BABYBEAR: np.uint32}


def ntt_root(field, log_n):
    """psi_N, the primitive 2N-th root of unity the transform uses (rule in csrc/ntt.cuh)."""
    o = C.c_uint64()
    rc = lib().lf_ntt_root(field, log_n, C.byref(o))
    if rc:
        raise LfError(rc, lib().lf_last_error(None).decode())
    return int(o.value)


class NttPlan:
    """Batched negacyclic NTT over Z_p[X]/(X^N + 1) (K12, include/lf_b200.h).  Arrays are (batch, N) of uint64 (Goldilocks)
    or uint32 (BabyBear), canonical values, natural order on both sides."""

    def __init__(self, ctx, field, log_n):
        self.ctx, self.field, self.log_n, self.n, self.dtype = ctx, field, log_n, 1 << log_n, NTT_DTYPE.get(field, np.uint64)
        h = vp(); ctx.check(ctx.L.lf_ntt_plan_create(ctx.h, field, log_n, C.byref(h))); self.h = h

    def close(self):
        if self.h and self.ctx.h:
            self.ctx.L.lf_ntt_plan_free(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _host(self, fn, a):
        a = np.ascontiguousarray(a, dtype=self.dtype); assert a.size % self.n == 0
        o = np.empty_like(a)
        self.ctx.check(fn(self.ctx.h, self.h, a.ctypes.data_as(vp), o.ctypes.data_as(vp), a.size // self.n)); return o

    def forward(self, a):
        return self._host(self.ctx.L.lf_ntt_forward_host, a)

    def inverse(self, a):
        return self._host(self.ctx.L.lf_ntt_inverse_host, a)

    def negacyclic_mul(self, a, b):
        a = np.ascontiguousarray(a, dtype=self.dtype); b = np.ascontiguousarray(b, dtype=self.dtype); assert a.shape == b.shape
        o = np.empty_like(a)
        self.ctx.check(self.ctx.L.lf_ntt_negacyclic_mul_host(self.ctx.h, self.h, a.ctypes.data_as(vp), b.ctypes.data_as(vp), o.ctypes.data_as(vp), a.size // self.n))
        return o

    # device-pointer entry points (torch tensors' data_ptr(), any allocation 16-byte aligned); asynchronous on ctx.stream()
    def forward_device(self, d_in, d_out, batch):
        self.ctx.check(self.ctx.L.lf_ntt_forward_device(self.ctx.h, self.h, vp(d_in), vp(d_out), batch))

    def inverse_device(self, d_in, d_out, batch):
        self.ctx.check(self.ctx.L.lf_ntt_inverse_device(self.ctx.h, self.h, vp(d_in), vp(d_out), batch))

    def pointwise_mul_device(self, d_a, d_b, d_out, batch):
        self.ctx.check(self.ctx.L.lf_ntt_pointwise_mul_device(self.ctx.h, self.h, vp(d_a), vp(d_b), vp(d_out), batch))


class AjtaiCommitmentScheme:
    """crates/latticefold/src/commitment/commitment_scheme.rs:17-114; the matrix lives in HBM."""

    def __init__(self, ctx, matrix):
        matrix = np.ascontiguousarray(matrix, dtype=np.uint64); self.ctx = ctx
        kappa, n = matrix.shape[0], matrix.shape[1]
        h = vp(); ctx.check(ctx.L.lf_ajtai_create(ctx.h, kappa, n, ptr(matrix), C.byref(h))); self.h = h

    def kappa(self):
        return self.ctx.L.lf_ajtai_kappa(self.h)

    def width(self):
        return self.ctx.L.lf_ajtai_width(self.h)

    def commit(self, f):        # commit / commit_ntt
        o = np.empty((self.kappa(), self.ctx.d), dtype=np.uint64); self.ctx.check(self.ctx.L.lf_commit(self.ctx.h, self.h, f.h, ptr(o))); return o

    def commit_batch(self, fs):
        arr = (vp * len(fs))(*[f.h for f in fs]); o = np.empty((len(fs), self.kappa(), self.ctx.d), dtype=np.uint64)
        self.ctx.check(self.ctx.L.lf_commit_batch(self.ctx.h, self.h, arr, len(fs), ptr(o))); return o

    def commit_coeff(self, f_coeff):    # commitment_scheme.rs:80-87
        o = np.empty((self.kappa(), self.ctx.d), dtype=np.uint64); self.ctx.check(self.ctx.L.lf_commit_coeff(self.ctx.h, self.h, f_coeff.h, ptr(o))); return o

    def decompose_and_commit_coeff(self, f_coeff, B, L):    # commitment_scheme.rs:89-101
        o = np.empty((self.kappa(), self.ctx.d), dtype=np.uint64); self.ctx.check(self.ctx.L.lf_decompose_and_commit_coeff(self.ctx.h, self.h, f_coeff.h, B, L, ptr(o))); return o

    def decompose_and_commit_ntt(self, w, B, L):            # commitment_scheme.rs:103-113
        o = np.empty((self.kappa(), self.ctx.d), dtype=np.uint64); self.ctx.check(self.ctx.L.lf_decompose_and_commit_ntt(self.ctx.h, self.h, w.h, B, L, ptr(o))); return o

    def commit_pieces(self, f_coeff, b, K):                 # decompose_witness + commit_witnesses, decomposition.rs:162-201
        o = np.empty((K, self.kappa(), self.ctx.d), dtype=np.uint64); self.ctx.check(self.ctx.L.lf_commit_pieces(self.ctx.h, self.h, f_coeff.h, b, K, ptr(o))); return o

    def __del__(self):
        try:
            if self.ctx.h:
                self.ctx.L.lf_ajtai_free(self.ctx.h, self.h)
        except Exception:
            pass


class SparseMatrix:
    def __init__(self, ctx, M):
        self.ctx = ctx; h = vp()
        ctx.check(ctx.L.lf_sparse_create(ctx.h, M["nrows"], M["ncols"], ptr(M["row_ptr"]), ptr(M["col"]), ptr(M["val"]), C.byref(h))); self.h = h

    def mat_vec_mul(self, z):   # arith/utils.rs:52-65
        h = vp(); self.ctx.check(self.ctx.L.lf_spmv(self.ctx.h, self.h, z.h, C.byref(h))); return DeviceVec(self.ctx, h)

    def __del__(self):
        try:
            if self.ctx.h:
                self.ctx.L.lf_sparse_free(self.ctx.h, self.h)
        except Exception:
            pass


class MLSumcheck:
    """utils/sumcheck.rs:53-80 with the transcript on the host."""

    @staticmethod
    def prove_as_subprotocol(ctx, transcript, mles, nvars, degree, comb, want_final=False):
        d, tau = ctx.d, ctx.tau
        cs = Comb(); keep = []
        if comb["kind"] == "fold":
            cs.kind, cs.n_mu, cs.b = LF_COMB_FOLD, comb["mu"].shape[0], comb["b"]
            mu = np.ascontiguousarray(comb["mu"], dtype=np.uint64); keep.append(mu); cs.mu_host = ptr(mu)
        else:
            cs.kind = LF_COMB_LIN if comb["kind"] == "lin" else LF_COMB_PRODUCTS
            coef = np.ascontiguousarray(comb["coef"], dtype=np.uint64)
            idx = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.int32) for x in comb["idx"]]))
            idx_len = np.ascontiguousarray(np.array([len(x) for x in comb["idx"]], dtype=np.int32))
            keep += [coef, idx, idx_len]
            cs.n_terms, cs.coef_host, cs.idx, cs.idx_len = len(comb["idx"]), ptr(coef), idx.ctypes.data_as(i32p), idx_len.ctypes.data_as(i32p)
        arr = (vp * len(mles))(*[m.h for m in mles]); sc = vp()
        ctx.check(ctx.L.lf_sumcheck_begin(ctx.h, arr, len(mles), nvars, degree, C.byref(cs), C.byref(sc)))
        for m in mles:
            m.h = None   # ownership moved into the sumcheck
        try:
            ring_nv = np.zeros(d, dtype=np.uint64); ring_nv[::tau] = nvars
            ring_deg = np.zeros(d, dtype=np.uint64); ring_deg[::tau] = degree
            transcript.absorb(ring_nv); transcript.absorb(ring_deg)
            msgs = np.empty((nvars, degree + 1, d), dtype=np.uint64); point = np.empty((nvars, tau), dtype=np.uint64)
            prev = None
            for i in range(nvars):
                ctx.check(ctx.L.lf_sumcheck_round(sc, ptr(prev), ptr(msgs[i])))
                transcript.absorb(msgs[i])
                r = transcript.get_challenge()
                transcript.absorb(np.ascontiguousarray(np.broadcast_to(r, (ctx.S, tau)).reshape(d)))
                point[i] = r; prev = np.ascontiguousarray(r)
            final = None
            if want_final:
                final = np.empty((len(mles), d), dtype=np.uint64); ctx.check(ctx.L.lf_sumcheck_finish(sc, ptr(prev), ptr(final)))
        finally:
            ctx.L.lf_sumcheck_free(sc)
        return (msgs, point, final) if want_final else (msgs, point)


def nifs_verify(prob, transcript, proof):
    """NIFSVerifier::verify (nifs.rs:117-162) through the product library's host verifier: returns the folded LCCCS words, raises
    LfError on rejection.  Needs no GPU and none of the witness-sized inputs (A, w_i_f, w_acc_f may be absent from prob)."""
    L = lib()
    light = {k: v for k, v in prob.items() if k not in ("A", "w_i_f", "w_acc_f")}
    P, keep = make_problem(light)
    lc = np.empty(int(L.lf_lcccs_words(C.byref(P))), dtype=np.uint64)
    proof = np.ascontiguousarray(proof, dtype=np.uint64)
    rc = L.lf_nifs_verify(C.byref(P), transcript.h, ptr(proof), ptr(lc))
    if rc:
        raise LfError(rc, L.lf_last_error(None).decode())
    return lc


def proof_to_bytes(prob, proof):
    """LFProof::serialize_with_mode(Compress::Yes) of a flat u64 proof (nifs.rs:28-34; examples/e2e.rs:126-146)"""
    L = lib(); P, keep = make_problem({k: v for k, v in prob.items() if k not in ("A", "w_i_f", "w_acc_f")})
    proof = np.ascontiguousarray(proof, dtype=np.uint64); out = np.empty(int(L.lf_proof_wire_bytes(C.byref(P))), dtype=np.uint8)
    rc = L.lf_proof_serialize(C.byref(P), ptr(proof), out.ctypes.data_as(C.POINTER(C.c_uint8)))
    if rc:
        raise LfError(rc, L.lf_last_error(None).decode())
    return out.tobytes()


def proof_from_bytes(prob, data):
    """the inverse, validating length prefixes and canonical field elements like ark-serialize's deserializer"""
    L = lib(); P, keep = make_problem({k: v for k, v in prob.items() if k not in ("A", "w_i_f", "w_acc_f")})
    buf = np.frombuffer(bytes(data), dtype=np.uint8); out = np.empty(int(L.lf_proof_words(C.byref(P))), dtype=np.uint64)
    rc = L.lf_proof_deserialize(C.byref(P), buf.ctypes.data_as(C.POINTER(C.c_uint8)), buf.size, ptr(out))
    if rc:
        raise LfError(rc, L.lf_last_error(None).decode())
    return out


def linearization_verify(prob, transcript, lin_proof):
    """LFLinearizationVerifier::verify (nifs/linearization.rs:192-285), host code: returns the LCCCS words, raises LfError on rejection."""
    L = lib()
    light = {k: v for k, v in prob.items() if k not in ("A", "w_i_f", "w_acc_f", "acc")}
    P, keep = make_problem(light)
    lc = np.empty(int(L.lf_lcccs_words(C.byref(P))), dtype=np.uint64)
    lin_proof = np.ascontiguousarray(lin_proof, dtype=np.uint64)
    rc = L.lf_linearization_verify(C.byref(P), transcript.h, ptr(lin_proof), ptr(lc))
    if rc:
        raise LfError(rc, L.lf_last_error(None).decode())
    return lc


class NIFSProver:
    """NIFSProver::prove (crates/latticefold/src/nifs.rs:48-103).  The Ajtai matrix and the CCS are uploaded once."""

    def __init__(self, ctx, prob):
        self.ctx = ctx
        P, keep = make_problem(prob)
        h = vp(); ctx.check(ctx.L.lf_prover_create(ctx.h, C.byref(P), C.byref(h))); self.h = h
        self.proof_words = int(ctx.L.lf_proof_words(C.byref(P))); self.lcccs_words = int(ctx.L.lf_lcccs_words(C.byref(P)))
        self.n = prob["n"] // ctx.world      # witness elements held by this rank
        self.kappa = prob["kappa"]

    def close(self):
        if self.h:
            self.ctx.L.lf_prover_free(self.h); self.h = None

    def linearize(self, prob, transcript):
        P, keep = make_problem(prob); ccs = prob["ccs"]
        lc = np.empty(self.lcccs_words, dtype=np.uint64)
        pf = np.empty((ccs["s"] * (ccs["d"] + 2) + self.ctx.tau + ccs["t"]) * self.ctx.d, dtype=np.uint64)
        self.ctx.check(self.ctx.L.lf_linearize(self.h, C.byref(P), transcript.h, ptr(lc), ptr(pf))); return lc, pf

    def lin_proof_words(self, prob):
        ccs = prob["ccs"]; return (ccs["s"] * (ccs["d"] + 2) + self.ctx.tau + ccs["t"]) * self.ctx.d

    def linearize_resident(self, prob, w_i, transcript, out=None):
        """LFLinearizationProver::prove on a resident witness (BASELINE configs[2])"""
        P, keep = make_problem({k: v for k, v in prob.items() if k not in ("w_i_f", "w_acc_f")})
        lc, pf = out if out is not None else (np.empty(self.lcccs_words, dtype=np.uint64), np.empty(self.lin_proof_words(prob), dtype=np.uint64))
        self.ctx.check(self.ctx.L.lf_linearize_resident(self.h, C.byref(P), w_i, transcript.h, ptr(lc), ptr(pf))); return lc, pf

    def witness_commit(self, w, out=None):
        """Witness::commit (arith.rs:357-362): A f of a resident witness"""
        o = out if out is not None else np.empty((self.kappa, self.ctx.d), dtype=np.uint64)
        self.ctx.check(self.ctx.L.lf_witness_commit(self.h, w, ptr(o))); return o

    def prove(self, prob, transcript, want_f=True, out=None):
        """host inputs -> (proof, folded LCCCS, folded witness f); H2D of both witnesses and D2H of the results inside."""
        P, keep = make_problem(prob)
        proof, lc, f = out if out is not None else (np.empty(self.proof_words, dtype=np.uint64), np.empty(self.lcccs_words, dtype=np.uint64),
                                                    np.empty((self.n, self.ctx.d), dtype=np.uint64) if want_f else None)
        self.ctx.check(self.ctx.L.lf_nifs_prove(self.h, C.byref(P), transcript.h, ptr(proof), ptr(lc), ptr(f))); return proof, lc, f

    def upload_witness(self, f):
        f = np.ascontiguousarray(f, dtype=np.uint64); h = vp(); self.ctx.check(self.ctx.L.lf_prover_upload_witness(self.h, ptr(f), C.byref(h))); return h

    def free_witness(self, w):
        self.ctx.L.lf_witness_free(self.h, w)

    def download_witness(self, w):
        o = np.empty((self.n, self.ctx.d), dtype=np.uint64); self.ctx.check(self.ctx.L.lf_witness_download_f(self.h, w, ptr(o))); return o

    def prove_resident(self, prob, w_acc, w_i, transcript, keep_witness=False):
        P, keep = make_problem(prob)
        proof = np.empty(self.proof_words, dtype=np.uint64); lc = np.empty(self.lcccs_words, dtype=np.uint64); w = vp()
        self.ctx.check(self.ctx.L.lf_nifs_prove_resident(self.h, C.byref(P), w_acc, w_i, transcript.h, ptr(proof), ptr(lc), C.byref(w) if keep_witness else None))
        return (proof, lc, w) if keep_witness else (proof, lc)

    def timing_detail(self):
        buf = C.create_string_buffer(1 << 14); self.ctx.L.lf_prover_timing_detail(self.h, buf, len(buf))
        return [(l.split()[0], float(l.split()[1])) for l in buf.value.decode().splitlines()]

    def timings(self):
        t = (C.c_double * 5)(); self.ctx.L.lf_prover_last_timings(self.h, t)
        return dict(linearization_ms=t[0], decomposition_ms=t[1], folding_ms=t[2], host_transcript_ms=t[3], total_ms=t[4])
