"""CPU-side checks of the product library: it loads, exports every symbol include/lf_b200.h declares, its host-side
transcript / RotSum reproduce the reference KATs, and compute entry points fail loudly without a GPU (no fallback)."""
import json
import os
import re

import numpy as np
import pytest

import latticefold_b200 as lf
from latticefold_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
G = synth.RING_GOLDILOCKS


def test_library_exports_every_declared_symbol():
    L = lf.lib()
    hdr = open(os.path.join(ROOT, "include", "lf_b200.h")).read()
    declared = set(re.findall(r"\b(lf_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"lf_status"}
    assert declared == set(api.SYMBOLS), declared ^ set(api.SYMBOLS)
    for s in declared:
        assert hasattr(L, s), s


def test_product_transcript_kats():
    # crates/latticefold/src/transcript/poseidon.rs:86-142, through the product's own Poseidon
    g = json.load(open(os.path.join(GOLD, "transcript_goldilocks.json")))
    t = lf.Transcript(G); t.absorb_base(np.array(g["absorbed"], dtype=np.uint64))
    assert list(map(int, t.get_challenge())) == g["big_challenge"]
    t = lf.Transcript(G); t.absorb_base(np.array(g["absorbed"], dtype=np.uint64))
    assert list(map(int, t.get_short_challenge())) == g["small_challenge_coeffs"]


def test_product_transcript_matches_oracle_on_long_schedule(oracle):
    # same absorb / challenge schedule on both implementations, crossing many rate boundaries
    a, b = lf.Transcript(G), oracle.transcript(G)
    els = synth.uniform_field(synth.RINGS[G]["p"], 37 * 24, 5).reshape(37, 24)
    for step in range(6):
        a.absorb(els[step * 5:(step + 1) * 5 + step]); b.absorb(els[step * 5:(step + 1) * 5 + step])
        a.absorb_tag("beta_s"); b.absorb_tag("beta_s")
        for _ in range(step + 1):
            assert np.array_equal(a.get_challenge(), b.get_challenge())
        assert np.array_equal(a.get_short_challenge(), b.get_short_challenge())


def test_product_rot_lin_combination_kat():
    # crates/cyclotomic-rings/src/rotation.rs:174-776
    g = json.load(open(os.path.join(GOLD, "rotsum_goldilocks.json")))
    rho = np.array(g["rho"], dtype=np.uint64).reshape(3, 24); theta = np.array(g["theta"], dtype=np.uint64).reshape(3, 3, 24)
    out = np.empty((3, 24), dtype=np.uint64)
    rc = lf.lib().lf_rot_lin_combination(G, api.ptr(rho), api.ptr(theta), 3, api.ptr(out))
    assert rc == 0 and np.array_equal(out, np.array(g["expected"], dtype=np.uint64).reshape(3, 24))


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lf.LfError) as e:
        lf.Context(G, 0)
    assert e.value.code == -20   # LF_ERR_CUDA


def test_unsupported_ring_is_reported():
    h = api.vp()
    rc = lf.lib().lf_transcript_create(3, api.C.byref(h))       # the Stark ring (256-bit field) is not implemented
    assert rc == -8


@pytest.mark.parametrize("ring", [synth.RING_GOLDILOCKS, synth.RING_BABYBEAR, synth.RING_FROG])
def test_product_host_side_matches_oracle_all_rings(oracle, ring):
    """transcript (incl. the ring's short-challenge set) and RotSum of the product's host code vs the oracle"""
    R = synth.RINGS[ring]; d, tau = R["d"], R["tau"]
    a, b = lf.Transcript(ring), oracle.transcript(ring)
    els = synth.uniform_field(R["p"], 9 * d, 5).reshape(9, d)
    for step in range(3):
        a.absorb(els[step * 3:(step + 1) * 3]); b.absorb(els[step * 3:(step + 1) * 3])
        a.absorb_tag("rho_s"); b.absorb_tag("rho_s")
        assert np.array_equal(a.get_challenge(), b.get_challenge())
        assert np.array_equal(a.get_short_challenge(), b.get_short_challenge())
    rho = synth.uniform_field(R["p"], 2 * d, 6).reshape(2, d); theta = synth.uniform_field(R["p"], 2 * tau * d, 7).reshape(2, tau, d)
    out = np.empty((tau, d), dtype=np.uint64)
    assert lf.lib().lf_rot_lin_combination(ring, api.ptr(rho), api.ptr(theta), 2, api.ptr(out)) == 0
    assert np.array_equal(out, oracle.rot_lin_combination(ring, rho, theta))
