"""CPU: the host-side lowering (BN folding, head fusion, residual / upsample wiring) reproduces the forward oracle."""
import numpy as np
import pytest
import torch

import cases
from centernet_lightning_b200 import plan as P
from oracle import spec_model
from plan_emulator import run_plan


@pytest.fixture(scope="module")
def det():
    m = spec_model.synth_init(spec_model.build_spec_model(80), seed=0)
    return m, P.build_plan(m.state_dict())


def test_macs_match_survey(det):
    _, pl = det
    assert abs(pl.macs(512, 512) / 1e9 - 90.660) < 0.01            # SURVEY 8d / BASELINE.md section 2
    assert pl.macs(1024, 1024) == 4 * pl.macs(512, 512)
    assert len(pl.ops) == 50


def test_plan_reproduces_oracle_fp32(det):
    m, pl = det
    x = cases.make_image(cases.FORWARD_CASES["det64"])
    with torch.no_grad():
        ref = m(x)
    out = run_plan(pl, x)
    for k in ref:
        np.testing.assert_allclose(out[k].numpy(), ref[k].numpy(), rtol=0, atol=2e-4)


def test_tracking_plan_has_three_heads():
    m = spec_model.synth_init(spec_model.build_spec_model(2, reid_dim=64), seed=1)
    pl = P.build_plan(m.state_dict(), head_names=("heatmap", "box_2d", "reid"))
    assert abs(pl.macs(512, 512) / 1e9 - 119.592) < 0.01
    x = cases.make_image(cases.FORWARD_CASES["track64"])
    with torch.no_grad():
        ref = m(x)
    out = run_plan(pl, x)
    assert set(out) == {"heatmap", "box_2d", "reid"}
    for k in ref:
        np.testing.assert_allclose(out[k].numpy(), ref[k].numpy(), rtol=0, atol=2e-4)


def test_split_precision_meets_parity_bar_and_single_pass_does_not(det):
    """Why CNL_PRECISION_SPLIT is the default: fp16 hi+lo operands (3 tensor-core passes, fp32 accumulate)
    keep the head outputs within 1e-3 of fp32; one fp16 pass does not."""
    m, pl = det
    x = cases.make_image(dict(n=1, size=128, img_seed=3))
    with torch.no_grad():
        ref = m(x)
    split = run_plan(pl, x, act_dtype=torch.float16, split=True)
    single = run_plan(pl, x, act_dtype=torch.float16)
    for k in ref:
        assert (split[k] - ref[k]).abs().max() < 1e-3
    assert max((single[k] - ref[k]).abs().max() for k in ref) > 1e-3
