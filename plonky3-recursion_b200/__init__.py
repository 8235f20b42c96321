"""plonky3-recursion_b200 — B200-native batch-STARK prover for Plonky3-recursion's prove_next_layer hot path.

Import with `importlib.import_module("plonky3-recursion_b200")` (the directory name carries a hyphen).
The product path is the CUDA library built from csrc/ (see lib.py); there is no CPU fallback.
"""
from . import abi, air, field, poseidon2_params, symbolic  # noqa: F401
