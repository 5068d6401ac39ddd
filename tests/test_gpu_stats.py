"""GPU parity of the validation feature statistics against the reference's topk_feature_extraction golden."""
from collections import namedtuple

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_feature_statistics_match_reference():
    from freud_b200.utils.validation import l1_feature_extraction, topk_feature_extraction
    from tests.util import GOLDEN

    z = np.load(f"{GOLDEN}/feature_stats.npz")
    Enc = namedtuple("Enc", "top_acts top_indices")
    Out = namedtuple("Out", "encoded")
    out = Out(Enc(torch.from_numpy(z["acts"]).cuda(), torch.from_numpy(z["idx"]).cuda()))
    got = topk_feature_extraction(out, 96, 1, "cuda")
    assert np.array_equal(got.cpu().numpy(), z["topk_max"])
    L = namedtuple("L", "latent")
    got = l1_feature_extraction(Out(L(torch.from_numpy(z["latent"]).cuda())))
    assert np.array_equal(got.cpu().numpy(), z["l1_max"])
