"""The torch generator of bench-size sketches gives the numpy generator's sketches element for element."""
import numpy as np
import pytest
import torch

from public_kssd_b200 import synth


@pytest.mark.parametrize("n,c,cs,ms", [(45, 500, 9, None), (70, 400, 7, 1003), (23, 64, 20, None), (41, 3000, 2, 7)])
def test_torch_generator_matches_numpy(n, c, cs, ms):
    want_c, want_i = synth.synth_sketches(n, c, seed=5, cluster_size=cs, member_seed=ms)
    got_c, got_i = synth.synth_sketches_torch(n, c, seed=5, device=torch.device("cpu"), cluster_size=cs, member_seed=ms, block_clusters=3)
    assert np.array_equal(got_i.numpy().astype(np.uint64), want_i)
    assert np.array_equal(got_c.numpy().view(np.uint32), want_c)


def test_torch_generator_handles_repeats_inside_a_sketch():
    # 12 code bits, 300 codes per genome: ancestors and members repeat codes, np.unique shortens them
    want_c, want_i = synth.synth_sketches(30, 300, seed=2, cluster_size=5, code_bits=12)
    got_c, got_i = synth.synth_sketches_torch(30, 300, seed=2, device=torch.device("cpu"), cluster_size=5, code_bits=12)
    assert np.array_equal(got_i.numpy().astype(np.uint64), want_i)
    assert np.array_equal(got_c.numpy().view(np.uint32), want_c)
