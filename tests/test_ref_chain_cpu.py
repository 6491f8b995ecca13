"""On whatever CPU runs the suite, torch's own ops (MKL SGEMM, ATen CPU trilinear) must agree with the
oracle's stated roundings — the same check make_golden.py did against the reference in the build container."""
import numpy as np
import pytest
import torch

from conftest import load_case
from ref_chain import torch_chain_indices


@pytest.mark.parametrize("name", ["tiny_s0", "tiny_sp15", "small_mh"])
def test_oracle_matches_torch_cpu_ops(oracle, name):
    sh, x, g = load_case(name)
    ref = torch_chain_indices(x["c3"], x["chm"], x["cam"], x["intr"], x["dist"], sh.G, sh.spacing, sh.hs, "cpu").numpy()
    mine = oracle.reproject_indices(x["c3"], x["chm"], x["cam"], x["intr"], x["dist"], sh.G, sh.spacing, sh.hs)
    assert np.array_equal(ref, mine)
    if "idx" in g:
        assert np.array_equal(ref, g["idx"])
