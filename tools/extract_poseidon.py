#!/usr/bin/env python3
"""Extract the Poseidon round constants / MDS matrices (numeric data only) from the
reference tree into compact binary tables.

Source of the numbers: /root/reference/crates/cyclotomic-rings/src/rings/poseidon/{goldilocks,babybear,frog}.rs
(`Fq::from(0x..._i128)` literals; `PoseidonConfig::new(full, partial, alpha, mds, ark, rate, capacity)`).
Output: latticefold_b200/data/poseidon_<ring>.bin
  header  : 6 x u64 LE  = full_rounds, partial_rounds, alpha, rate, capacity, width
  ark     : (full+partial) x width u64 LE, raw literals (NOT reduced; loaders reduce mod p)
  mds     : width x width u64 LE, raw literals
Run once in the build container (the reference is not present on the GPU box); the .bin files are committed.
"""
import re, struct, sys, pathlib

REF = pathlib.Path("/root/reference/crates/cyclotomic-rings/src/rings/poseidon")
OUT = pathlib.Path(__file__).resolve().parent.parent / "latticefold_b200" / "data"

def parse(path):
    src = path.read_text()
    full = int(re.search(r"full_rounds\s*=\s*(\d+)", src).group(1))
    partial = int(re.search(r"partial_rounds\s*=\s*(\d+)", src).group(1))
    alpha = int(re.search(r"alpha\s*=\s*(\d+)", src).group(1))
    m = re.search(r"PoseidonConfig::<Fq>::new\([^)]*?,\s*(\d+),\s*(\d+)\)", src)
    rate, cap = int(m.group(1)), int(m.group(2))
    i_ark = src.index("let ark")
    i_mds = src.index("let mds")
    lit = re.compile(r"Fq::from\(\s*(0x[0-9a-fA-F_]+?)_i128\s*\)")
    first, second = (i_ark, i_mds) if i_ark < i_mds else (i_mds, i_ark)
    a = [int(x.replace("_", ""), 16) for x in lit.findall(src[first:second])]
    b = [int(x.replace("_", ""), 16) for x in lit.findall(src[second:])]
    ark, mds = (a, b) if i_ark < i_mds else (b, a)
    width = rate + cap
    assert len(ark) == (full + partial) * width, (len(ark), full, partial, width)
    assert len(mds) == width * width, len(mds)
    assert all(0 <= v < 2**64 for v in ark + mds)
    return full, partial, alpha, rate, cap, width, ark, mds

def main():
    OUT.mkdir(parents=True, exist_ok=True)
    for ring in ("goldilocks", "babybear", "frog"):
        full, partial, alpha, rate, cap, width, ark, mds = parse(REF / f"{ring}.rs")
        blob = struct.pack("<6Q", full, partial, alpha, rate, cap, width)
        blob += struct.pack(f"<{len(ark)}Q", *ark) + struct.pack(f"<{len(mds)}Q", *mds)
        (OUT / f"poseidon_{ring}.bin").write_bytes(blob)
        print(ring, full, partial, alpha, rate, cap, width, len(blob), "bytes")

if __name__ == "__main__":
    main()
