"""CPU-only: the oracle restatement reproduces the golden outputs generated from the reference itself."""
import os

import torch

from conftest import rel_err


def test_encoder_oracle_matches_reference_golden(golden_dir):
    from oracle.encoder_oracle import encoder_forward, random_encoder_state
    gold = torch.load(os.path.join(golden_dir, "encoder.pt"))
    for name in ("img64_s7", "img256_s7", "img256_s14"):
        g = gold[name]
        sd = random_encoder_state(seed=g["weight_seed"], layer_scale=1.0)
        x = torch.randn(*g["shape"], generator=torch.Generator().manual_seed(g["input_seed"]))
        y = encoder_forward(sd, x, g["enc_size"])
        assert y.shape == g["out"].shape
        assert rel_err(y, g["out"]) < 1e-5, name


def test_adaptive_pool_8_to_7_is_avgpool_2_1():
    # SURVEY.md §8c invariant (i)
    import torch.nn.functional as F
    x = torch.randn(2, 16, 8, 8)
    assert torch.equal(F.adaptive_avg_pool2d(x, (7, 7)), F.avg_pool2d(x, 2, 1))
