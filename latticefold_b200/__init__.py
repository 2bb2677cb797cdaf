"""latticefold_b200: Blackwell-native LatticeFold prover hot path.

Importing the package does not load CUDA; `latticefold_b200.lib()` loads the C-ABI shared library
(latticefold_b200/_lib/liblf_b200.so) and raises if it is missing -- there is no CPU fallback.
"""
from . import synth  # noqa: F401
from .api import (AjtaiCommitmentScheme, Context, DeviceVec, LfError, MLSumcheck, NIFSProver, NttPlan, SparseMatrix,  # noqa: E402,F401
                  Transcript, lib, linearization_verify, nifs_verify, ntt_root, proof_from_bytes, proof_to_bytes)
