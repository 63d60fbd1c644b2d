"""bambu_b200 -- B200-native implementation of bambu's quantification hot path
(the per-gene-group EM of ``src/em.cpp``, reference bambu 3.3.5).

Only what the path needs lives here: the arena format, the host-side mirror of
the reference's quantification interface, the synthetic workload generators,
and ``csrc/`` (CUDA kernels + the C-ABI shared library)."""
from .arena import Arena  # noqa: F401

__version__ = "0.1.0"
