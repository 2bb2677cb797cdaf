"""GPU parity of the batched negacyclic NTT (K12, csrc/ntt.cu) against the CPU oracle, through the C ABI."""
import numpy as np
import pytest

import latticefold_b200 as lf
from latticefold_b200 import synth
from latticefold_b200.api import FIELD_BABYBEAR, FIELD_GOLDILOCKS

pytestmark = pytest.mark.gpu
P = {0: 0xFFFFFFFF00000001, 1: 2013265921}


@pytest.fixture(scope="module")
def ctx():
    c = lf.Context(synth.RING_GOLDILOCKS, 0)
    yield c
    c.close()


def rnd(field, shape, seed, edge=True):
    a = synth.uniform_field(P[field], int(np.prod(shape)), seed).reshape(shape)
    if edge:      # extreme residues in the first rows
        a.reshape(-1)[:4] = [0, P[field] - 1, 1, P[field] - 2]
    return a


@pytest.mark.parametrize("field", [FIELD_GOLDILOCKS, FIELD_BABYBEAR])
@pytest.mark.parametrize("log_n", list(range(8, 17)))
def test_forward_inverse_match_oracle(ctx, oracle, field, log_n):
    batch = 5 if log_n <= 12 else 3            # ragged against the polynomials-per-CTA grouping of the small sizes
    a = rnd(field, (batch, 1 << log_n), 100 + log_n)
    plan = lf.NttPlan(ctx, field, log_n)
    f = plan.forward(a)
    assert f.dtype == plan.dtype and np.array_equal(f.astype(np.uint64), oracle.ntt(field, log_n, a))
    i = plan.inverse(f)
    assert np.array_equal(i.astype(np.uint64), a)
    g = rnd(field, (batch, 1 << log_n), 200 + log_n)
    assert np.array_equal(plan.inverse(g).astype(np.uint64), oracle.ntt(field, log_n, g, inverse=True))
    plan.close()


@pytest.mark.parametrize("field", [FIELD_GOLDILOCKS, FIELD_BABYBEAR])
def test_definition_small(ctx, oracle, field):
    a = rnd(field, (1, 256), 7)
    plan = lf.NttPlan(ctx, field, 8)
    assert np.array_equal(plan.forward(a)[0].astype(np.uint64), oracle.ntt_naive(field, 8, a[0]))
    plan.close()


@pytest.mark.parametrize("field", [FIELD_GOLDILOCKS, FIELD_BABYBEAR])
@pytest.mark.parametrize("log_n", [8, 12, 16])
def test_negacyclic_product(ctx, oracle, field, log_n):
    n = 1 << log_n
    a, b = rnd(field, (2, n), 31), rnd(field, (2, n), 32)
    b[1] = 0; b[1, 1] = 1          # times X: a negacyclic rotation with a sign flip
    plan = lf.NttPlan(ctx, field, log_n)
    c = plan.negacyclic_mul(a, b).astype(np.uint64)
    if log_n <= 12:
        assert np.array_equal(c[0], oracle.ntt_schoolbook(field, log_n, a[0], b[0]))
    else:           # schoolbook at 2^16 is minutes on the CPU: check through the oracle's transform instead
        fa, fb = oracle.ntt(field, log_n, a[:1]), oracle.ntt(field, log_n, b[:1])
        prod = (fa.astype(object) * fb.astype(object) % P[field]).astype(np.uint64)
        assert np.array_equal(c[0], oracle.ntt(field, log_n, prod, inverse=True)[0])
    rot = np.concatenate([[(P[field] - int(a[1, -1])) % P[field]], a[1, :-1]]).astype(np.uint64)
    assert np.array_equal(c[1], rot)
    plan.close()


@pytest.mark.parametrize("field,log_n,batch", [(FIELD_GOLDILOCKS, 12, 4096), (FIELD_GOLDILOCKS, 16, 300), (FIELD_BABYBEAR, 10, 20000), (FIELD_BABYBEAR, 16, 700)])
def test_full_size_properties(ctx, field, log_n, batch):
    """sizes the oracle would take minutes on: round trip, linearity and the transform of a monomial, on device buffers"""
    import torch
    n = 1 << log_n; p = P[field]
    tdt = torch.int64 if field == FIELD_GOLDILOCKS else torch.int32
    a = rnd(field, (batch, n), 5, edge=False)
    plan = lf.NttPlan(ctx, field, log_n)
    da = torch.from_numpy(a.astype(plan.dtype).view(np.int64 if field == 0 else np.int32)).cuda()
    df = torch.empty_like(da); db = torch.empty_like(da)
    torch.cuda.synchronize()
    plan.forward_device(da.data_ptr(), df.data_ptr(), batch)
    plan.inverse_device(df.data_ptr(), db.data_ptr(), batch)
    ctx.sync()
    assert torch.equal(da, db)
    # linearity on rows 0, 1: NTT(a0 + a1) = NTT(a0) + NTT(a1)
    f = df[:2].cpu().numpy().view(plan.dtype).astype(object)
    s = ((a[0].astype(object) + a[1].astype(object)) % p).astype(np.uint64)
    fs = plan.forward(s.reshape(1, n))[0].astype(object)
    assert np.array_equal(fs, (f[0] + f[1]) % p)
    # in place on a chunk that spans several L2-sized scratch chunks of the four-step path
    plan.forward_device(da.data_ptr(), da.data_ptr(), batch); ctx.sync()
    assert torch.equal(da, df)
    plan.close()


def test_errors(ctx):
    with pytest.raises(lf.LfError) as e:
        lf.NttPlan(ctx, FIELD_GOLDILOCKS, 17)
    assert e.value.code == -8
    with pytest.raises(lf.LfError):
        lf.NttPlan(ctx, 5, 10)
