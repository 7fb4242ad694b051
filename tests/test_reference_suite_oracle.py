"""The reference's test-suite run against the oracle (CPU)."""
import scipy.sparse.linalg as spla

import oracle
import reference_suite as rs


def test_matrix_vector_product():
    rs.matrix_vector_product(oracle)


def test_explicit_implementation_blockdiagonal_preconditioner():
    rs.explicit_blockdiagonal_preconditioner(oracle)


def test_preconditioner_times_matrix_gives_identity():
    rs.preconditioner_times_matrix_gives_identity(oracle)


def test_block_diagonal_operator():
    rs.block_diagonal_operator(oracle)


def test_SPD_properties_block_diagonal_preconditioner():
    rs.spd_properties_block_diagonal_preconditioner(oracle)


def test_toeplitz_vector_products():
    rs.toeplitz_vector_products(oracle)


def test_deflation_operator():
    rs.deflation_operator(oracle)


def test_coarse_operator():
    rs.coarse_operator(oracle)


def test_2level_preconditioner():
    rs.two_level_preconditioner(oracle, spla.cg)
