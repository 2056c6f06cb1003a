"""zkp_b200 -- B200-native ristretto255 multi-scalar-multiplication engine behind the call sites of
dalek-cryptography/zkp's toolbox (include/zkp_b200.h).

The package holds only what the hot path needs: `csrc/` (CUDA kernels + the C ABI), `native`/`engine`
(ctypes face of the C ABI) and `toolbox` (host-side mirror of the reference's Prover / Verifier /
BatchVerifier that routes every MSM, compress and decompress through the engine).  Nothing here imports
`oracle/`; without the CUDA library the package raises.
"""
from .native import NativeLibraryMissing, load as load_native  # noqa: F401
from .engine import Engine, EngineError  # noqa: F401
