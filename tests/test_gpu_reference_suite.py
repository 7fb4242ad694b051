"""The reference's test-suite run against the CUDA product through its drop-in operators (B200)."""
import pytest
import scipy.sparse.linalg as spla

import reference_suite as rs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cm():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import cosmomap2_b200
    return cosmomap2_b200


def test_matrix_vector_product(cm):
    rs.matrix_vector_product(cm)


def test_explicit_implementation_blockdiagonal_preconditioner(cm):
    rs.explicit_blockdiagonal_preconditioner(cm)


def test_preconditioner_times_matrix_gives_identity(cm):
    rs.preconditioner_times_matrix_gives_identity(cm)


def test_block_diagonal_operator(cm):
    rs.block_diagonal_operator(cm)


def test_SPD_properties_block_diagonal_preconditioner(cm):
    rs.spd_properties_block_diagonal_preconditioner(cm)


def test_toeplitz_vector_products(cm):
    rs.toeplitz_vector_products(cm)


def test_deflation_operator(cm):
    rs.deflation_operator(cm)


def test_coarse_operator(cm):
    rs.coarse_operator(cm)


def test_2level_preconditioner_scipy_driver(cm):
    rs.two_level_preconditioner(cm, spla.cg)


def test_2level_preconditioner_device_cg(cm):
    rs.two_level_preconditioner(cm, cm.cg)
