"""evosops_b200 — B200-native batched SOPS genome-fitness evaluation (the EvoSOPS SOPSCore hot path).

Layout: csrc/ (sm_100a CUDA kernels + the C ABI of include/evosops.h), capi.py (ctypes binding of that ABI),
host/ (C++ mirror of the reference's SOPSCore / GACore module API and CLI over the same ABI), sharding.py
(one-process-per-GPU partitioning of a population).  Nothing here imports oracle/."""
from .capi import (AGG, SEP, COAT, LOCO, RNG_VERIFY, RNG_FAST, F_DUMP, F_INIT_ONLY, F_THEORY_TABLE,  # noqa: F401
                   F_FORCE_WARP, F_FORCE_CTA, F_ROUNDS_SERIAL, F_ROUNDS_PARALLEL, F_ROUNDS_ADAPTIVE,
                   GENOME_LEN, GENOME_SHAPE, Context, SopsError, instance_shape, coat_distance_grid)

__all__ = ["AGG", "SEP", "COAT", "LOCO", "RNG_VERIFY", "RNG_FAST", "F_DUMP", "F_INIT_ONLY", "F_THEORY_TABLE", "F_FORCE_WARP", "F_FORCE_CTA",
           "F_ROUNDS_SERIAL", "F_ROUNDS_PARALLEL", "F_ROUNDS_ADAPTIVE",
           "GENOME_LEN", "GENOME_SHAPE", "Context", "SopsError", "instance_shape", "coat_distance_grid"]
