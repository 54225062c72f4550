"""CPU oracle for the matrix-free spectral embedding path of SnapATAC2.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline -- never as a fallback for the CUDA path.

PARITY PINNED to reference-executed code: the reference ships no golden vectors
or known-answer tests for this path and its Rust/PyO3
extension cannot be built in this image (no cargo/rustc), but the reference also
carries a pure-Python statement of the same algorithm.
``tests/golden/make_ref_golden.py`` loads ``tools/_embedding.py`` unmodified
(stub modules stand in for the two imports of the compiled extension) and runs
``SpectralMatrixFree(...).fit(X).transform()`` (:434-481), ``orthogonalize``
(:397-413) and the wrapper ``spectral`` (:129-295) on every fixture input, and
executes the Python snippet embedded in ``frobenius_norm`` (embedding.rs:456-460)
on both scipy containers; the outputs are committed as ``tests/golden/*_ref.npz``.
``tests/test_oracle.py`` asserts oracle == those vectors (1e-10), the ``-m gpu``
tests assert CUDA == those vectors at north_star's tolerances.  Additional pins:
an independent dense ``numpy.linalg.eigh`` statement (``dense_check``).  Not
executable here and therefore restatement-only: the Rust IDF closed form
(embedding.rs:269-286) and Rust's ``StdRng`` draws (Nystrom landmarks, the 2000-row
sample of multi_spectral).
"""

from .reference_restatement import (  # noqa: F401
    idf,
    normalize,
    spectral_mf,
    spectral_embedding,
    spectral,
    multi_spectral_embedding,
    multi_spectral,
    matrix_free_twin,
    dense_check,
    operator_pieces,
    compute_degrees,
    nystrom,
    orthogonalize,
    spectral_embedding_nystrom,
)
from . import knn_restatement as knn  # noqa: F401  (pp.knn: the consumer next to the path)
