"""The tensor-core digit commit (csrc/commit_mma.cuh: tcgen05.mma.kind::i8 on byte limbs of the Ajtai matrix x int8 digits) and
the commit entry points of commitment_scheme.rs:72-114, against the CPU oracle.  The same inputs are also run with the tensor
path disabled (LF_COMMIT_MMA=0: lazily reduced dot products), so a failure tells which of the two paths differs."""
import os

import numpy as np
import pytest

from latticefold_b200 import synth
from tests.helpers import rand_elems

pytestmark = pytest.mark.gpu
G = synth.RING_GOLDILOCKS
D = 24
P = synth.RINGS[G]["p"]


def small_coeff_vec(n, bits, seed):
    """coefficient-form vector with balanced coefficients in (-2^(bits-1), 2^(bits-1)) (what a B-bounded witness looks like)"""
    v = synth.splitmix64(seed, n * D).astype(np.int64) >> np.int64(64 - bits)          # arithmetic shift: signed
    return np.ascontiguousarray((v.astype(object) % P).astype(np.uint64).reshape(n, D))


def oracle_pieces(oracle, A, f_coeff, b, K):
    pcs = oracle.decompose_to_vec(G, f_coeff, b, K)
    return np.stack([oracle.commit(G, A, oracle.crt(G, pcs[k])) for k in range(K)])


@pytest.mark.parametrize("mma", ["1", "0"])
@pytest.mark.parametrize("kappa,n,K", [(1, 1, 1), (2, 64, 2), (3, 100, 5), (5, 1000, 16), (26, 4096, 15), (7, 5000, 17), (4, 70000, 3)])
def test_commit_pieces_matches_oracle(gpu, oracle, kappa, n, K, mma):
    old = os.environ.get("LF_COMMIT_MMA")
    os.environ["LF_COMMIT_MMA"] = mma
    try:
        ctx = gpu.Context(G, 0)
        A = rand_elems(G, kappa * n, 60 + kappa).reshape(kappa, n, D)
        f = small_coeff_vec(n, K, 61 + n)
        sch = gpu.AjtaiCommitmentScheme(ctx, A)
        got = sch.commit_pieces(ctx.upload(f, 1), 2, K)
        assert np.array_equal(got, oracle_pieces(oracle, A, f, 2, K))
        del sch
        ctx.close()
    finally:
        if old is None:
            os.environ.pop("LF_COMMIT_MMA", None)
        else:
            os.environ["LF_COMMIT_MMA"] = old


def test_commit_pieces_extreme_digits(gpu, oracle):
    """all digits +1 / all -1 / alternating, and matrix limbs 0 and p-1: the byte-limb sums at their extremes"""
    ctx = gpu.Context(G, 0)
    kappa, n, K = 3, 8192, 4
    A = rand_elems(G, kappa * n, 70).reshape(kappa, n, D)
    A[0] = P - 1; A[1, ::2] = 0; A[2, :, 1::2] = 0xFFFFFFFF
    sch = gpu.AjtaiCommitmentScheme(ctx, A)
    for val in (2 ** K - 1, P - (2 ** K - 1)):
        f = np.full((n, D), val, dtype=np.uint64)
        assert np.array_equal(sch.commit_pieces(ctx.upload(f, 1), 2, K), oracle_pieces(oracle, A, f, 2, K))
    f = np.full((n, D), 5, dtype=np.uint64); f[1::2] = P - 5
    assert np.array_equal(sch.commit_pieces(ctx.upload(f, 1), 2, K), oracle_pieces(oracle, A, f, 2, K))
    with pytest.raises(gpu.LfError) as e:          # a coefficient that needs more than K digits
        sch.commit_pieces(ctx.upload(np.full((n, D), 1 << K, dtype=np.uint64), 1), 2, K)
    assert e.value.code == -9
    with pytest.raises(gpu.LfError) as e:          # commitment_scheme.rs:37-44
        sch.commit_pieces(ctx.upload(f[:10], 1), 2, K)
    assert e.value.code == -1
    del sch
    ctx.close()


def test_commit_coeff_and_decompose_and_commit(gpu, oracle):
    """commit_coeff / decompose_and_commit_coeff / decompose_and_commit_ntt (commitment_scheme.rs:80-114)"""
    ctx = gpu.Context(G, 0)
    kappa, W, B, L = 4, 96, 1 << 16, 4
    n = W * L
    A = rand_elems(G, kappa * n, 80).reshape(kappa, n, D)
    sch = gpu.AjtaiCommitmentScheme(ctx, A)
    fc = rand_elems(G, n, 81)
    assert np.array_equal(sch.commit_coeff(ctx.upload(fc, 1)), oracle.commit(G, A, oracle.crt(G, fc)))
    w = rand_elems(G, W, 82)                                   # uniform coefficients need 4 digits of base 2^16
    want = oracle.commit(G, A, oracle.crt(G, oracle.gadget_decompose(G, w, B, L)))
    assert np.array_equal(sch.decompose_and_commit_coeff(ctx.upload(w, 1), B, L), want)
    wn = oracle.crt(G, w)
    assert np.array_equal(sch.decompose_and_commit_ntt(ctx.upload(wn), B, L), want)
    with pytest.raises(gpu.LfError) as e:
        sch.decompose_and_commit_coeff(ctx.upload(w[:10], 1), B, L)
    assert e.value.code == -1
    with pytest.raises(gpu.LfError) as e:
        sch.decompose_and_commit_coeff(ctx.upload(w, 1), 16, L)
    assert e.value.code == -9
    del sch
    ctx.close()
