"""GPU parity of the fused chains (cm2_amatvec_toeplitz, cm2_pointing_filter_mu) against golden vectors
produced by the reference's own operators (tests/golden/fused_chains.npz, make_golden.py::case_fused_chains),
with the assertion that the compositions really ran as the fused kernels."""
import pytest

import golden_cases as gc

pytestmark = pytest.mark.gpu


def test_fused_chains_against_reference_fixture():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import cosmomap2_b200 as cm
    from cosmomap2_b200 import linearoperators as lo
    gc.check_fused_chains(cm, expect_fused=lo)
