"""CPU oracle for the matrix-free spectral embedding path of SnapATAC2.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline -- never as a fallback for the CUDA path.

PARITY UNPINNED by the reference's own tests: the reference ships no golden
vectors or known-answer tests for this path, and its
Rust/PyO3 extension cannot be built in this image (no cargo/rustc).  The oracle
is therefore pinned by (1) being a line-by-line restatement of
``snapatac2-python/src/embedding.rs`` driving the *same* scipy ``eigsh`` call
the reference embeds verbatim (embedding.rs:158-171), (2) the reference's own
second statement of the algorithm, ``SpectralMatrixFree.fit`` / ``_eigen``
(tools/_embedding.py:447-481), restated in ``matrix_free_twin``, and (3) an
independent dense ``numpy.linalg.eigh`` cross-check (``dense_check``).
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
