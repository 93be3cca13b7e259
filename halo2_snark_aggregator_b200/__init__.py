"""halo2_snark_aggregator_b200 -- B200-native backend for the halo2 aggregation prover hot path.

Host-side mirror (Python, over the C ABI in include/h2agg.h) of the functions the reference's
`create_proof` call reaches (halo2-snark-aggregator-circuit/src/verify_circuit.rs:986-994):
    arithmetic.best_multiexp / best_fft      <- halo2_proofs::arithmetic
    domain.EvaluationDomain                  <- halo2_proofs::poly::EvaluationDomain
    params.ParamsKZG.commit_lagrange/commit  <- halo2_proofs::poly::kzg::commitment::ParamsKZG
The C++ mirror of the same surface is include/h2agg.hpp; the Rust binding is rust/h2agg-sys.
"""
from ._lib import H2aggError, LIB_PATH, declared_symbols, load  # noqa: F401
from .context import Context, default_context  # noqa: F401
from .arithmetic import best_fft, best_multiexp  # noqa: F401
from .domain import EvaluationDomain  # noqa: F401
from .params import ParamsKZG  # noqa: F401
from .witness import B200Context, B200EccChip, B200EncodeChip, B200ScalarChip  # noqa: F401
