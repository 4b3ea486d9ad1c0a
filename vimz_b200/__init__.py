"""vimz_b200 -- B200-native kernels for the per-step NIFS fold of zero-savvy/vimz's Nova prover.

`vimz_b200.nova` mirrors the nova-snark interface of that path on top of libvimz_gpu.so (C ABI in
include/vimz_gpu.h).  Importing this package loads the shared library and fails if it is missing.
"""
from ._lib import EXPORTED_SYMBOLS, LIB_PATH, InvalidIndex, InvalidWitnessLength, VimzError, lib  # noqa: F401
from .field import CURVES  # noqa: F401
from .nova import (  # noqa: F401
    CommitmentEngine,
    CommitmentKey,
    Engine,
    FoldAccumulator,
    NIFS,
    R1CSInstance,
    R1CSShape,
    R1CSWitness,
    RelaxedR1CSInstance,
    RelaxedR1CSWitness,
    TranscriptRO,
    UnSat,
    is_sat,
    is_sat_relaxed,
)

from .recursive import PublicParams, RecursiveSNARK, fold_input, verify_folded_proof  # noqa: F401
from .sonobe import SonobeNova  # noqa: F401

__version__ = "0.2.0"
