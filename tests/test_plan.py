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


@pytest.mark.parametrize("case", ["simple64", "deconv3_64", "deconv4_64"])
def test_simple_neck_plan_reproduces_oracle(case):
    """BASELINE configs[0] neck: conv + nearest x2 as an upsampling store; ConvTranspose2d(k=3/4, stride 2) as four
    sub-pixel phase convolutions (reference models/layers.py:86-99)."""
    kw = cases.FORWARD_CASES[case]
    m = spec_model.synth_init(spec_model.build_spec_model(**kw["model"]), seed=kw["seed"])
    pl = P.build_plan(m.state_dict(), neck="simple")
    x = cases.make_image(kw)
    with torch.no_grad():
        ref = m(x)
    out = run_plan(pl, x)
    for k in ref:
        assert out[k].shape == ref[k].shape
        np.testing.assert_allclose(out[k].numpy(), ref[k].numpy(), rtol=0, atol=2e-4)
    n_phase = sum(1 for op in pl.ops if op.dst_up == 2 and op.dst_phase >= 0)
    assert n_phase == (12 if "deconv" in case else 0)


def test_conv_transpose_phase_lowering_matches_torch():
    """lower_conv_transpose on its own: every phase window / padding against F.conv_transpose2d, k = 3 and 4."""
    import torch.nn.functional as F
    from centernet_lightning_b200.plan import Plan, lower_conv_transpose
    g = torch.Generator().manual_seed(0)
    for k in (3, 4):
        w = torch.randn((64, 64, k, k), generator=g) * 0.1
        x = torch.randn((2, 64, 5, 7), generator=g)
        op_pad = k % 2
        ref = F.conv_transpose2d(x, w, None, stride=2, padding=(k + op_pad) // 2 - 1, output_padding=op_pad)
        pl = Plan()
        pl.add_buffer("image", 3, 1, fp32_nchw=True)
        pl.add_buffer("src", 64, 2)
        pl.add_buffer("dst", 64, 1)
        pl.ops = lower_conv_transpose("up", "src", "dst", w, None, relu=False)
        pl.outputs = {"y": "dst"}
        out = run_plan(pl, torch.zeros(1), extra={"src": x})["y"]
        np.testing.assert_allclose(out.numpy(), ref.numpy(), rtol=0, atol=1e-5)


def test_load_reference_state_dict_with_key_map():
    """Checkpoint hook (SURVEY 8b "state-dict names to accept"): Lightning-style keys with foreign inner names are
    remapped by regex rules; training-only modules are dropped; a stray key still raises under strict."""
    from centernet_lightning_b200.model import CenterNet
    spec = spec_model.synth_init(spec_model.build_spec_model(3), seed=4)
    foreign = {}
    for k, v in spec.state_dict().items():
        k = "model." + k
        k = k.replace("model.backbone.conv1.", "model.backbone.stem.0.").replace("model.backbone.bn1.", "model.backbone.stem.1.")
        k = k.replace(".block_1.conv.", ".block_1.0.").replace(".block_1.bn.", ".block_1.1.")
        foreign[k] = v
    foreign["model.backbone.fc.weight"] = torch.zeros(10, 512)
    foreign["evaluator.count"] = torch.zeros(1)
    net = CenterNet(3)
    with pytest.raises(RuntimeError):
        net.load_state_dict(foreign)
    rules = {r"backbone\.stem\.0\.": "backbone.conv1.", r"backbone\.stem\.1\.": "backbone.bn1.",
             r"\.block_1\.0\.": ".block_1.conv.", r"\.block_1\.1\.": ".block_1.bn."}
    net.load_reference_state_dict(foreign, key_map=rules)
    got = net.state_dict()
    for k, v in spec.state_dict().items():
        assert torch.equal(got["model." + k], v), k
    foreign["model.neck.unknown.weight"] = torch.zeros(1)
    with pytest.raises(RuntimeError):
        net.load_reference_state_dict(foreign, key_map=rules)


def test_resnet50_bottleneck_plan_reproduces_oracle():
    """resnet50 (reference tests/test_models.py:37-39 lists it): Bottleneck blocks lowered to 1x1 / 3x3(stride) / 1x1+residual
    launches, FPN laterals on 256/512/1024/2048 channels; parameter count of the trunk = torchvision resnet50 minus fc."""
    m = spec_model.synth_init(spec_model.build_spec_model(4, backbone="resnet50"), seed=7)
    assert spec_model.count_params(m.backbone) == 23_508_032
    pl = P.build_plan(m.state_dict(), backbone="resnet50")
    x = cases.make_image(dict(n=1, size=64, img_seed=5))
    with torch.no_grad():
        ref = m(x)
    out = run_plan(pl, x)
    for k in ref:
        np.testing.assert_allclose(out[k].numpy(), ref[k].numpy(), rtol=0, atol=2e-4)


BREADTH = ["mbv2_fpn64", "mbv2_simple64", "r18_ida64", "r18_bifpn128", "mbv2_ida64_sep", "mbv2_bifpn64_sep"]


@pytest.mark.parametrize("case", BREADTH)
def test_breadth_plans_reproduce_oracle(case):
    """SURVEY 8f rank 4 remainder: MobileNetV2 (depthwise + pointwise stages, ReLU6, channels zero-padded to multiples of
    64), IDA / BiFPN necks lowered from the reference's Fuse node (models/layers.py:138-177) and make_conv's separable
    branch (:56-68): the op list reproduces the oracle graph in fp32."""
    kw = cases.FORWARD_CASES[case]
    m = spec_model.synth_init(spec_model.build_spec_model(**kw["model"]), seed=kw["seed"], **kw.get("init", {}))
    pl = P.build_plan(m.state_dict(), backbone=kw["model"].get("backbone", "resnet34"), neck=kw["model"].get("neck", "FPN"),
                      head_depth=kw["model"]["head_config"]["depth"])
    x = cases.make_image(kw)
    with torch.no_grad():
        ref = m(x)
    out = run_plan(pl, x)
    for k in ref:
        assert out[k].shape == ref[k].shape
        np.testing.assert_allclose(out[k].numpy(), ref[k].numpy(), rtol=0, atol=5e-4)
    kinds = {op.kind for op in pl.ops}
    if "mbv2" in case:
        assert {"stem3x3", "dw", "conv"} <= kinds
        assert all(b.channels % 64 == 0 for b in pl.buffers.values() if not b.fp32_nchw)
    if "bifpn" in case or "sep" in case:
        assert "fuse" in kinds
    # the hi+lo fp16 operand storage of the engine keeps these graphs inside the parity bar as well
    split = run_plan(pl, x, act_dtype=torch.float16, split=True)
    for k in ref:
        assert (split[k] - ref[k]).abs().max() < 1e-3


def test_fuse_lowering_folds_the_fusion_weights():
    """Weighted fusion (reference models/layers.py:167-171): out = sum_i relu(w_i) x_i / (sum_j relu(w_j) + 1e-6).  Projected
    inputs carry their weight inside the 1x1 conv; un-projected ones in the fuse op's scale; a negative weight is clamped."""
    torch.manual_seed(0)
    node = spec_model.Fuse([64, 128, 64], 64, "down", conv_type="separable", weighted_fusion=True).eval()
    with torch.no_grad():
        node.weights.copy_(torch.tensor([0.7, -0.3, 1.9]))
        for mod in node.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.1); mod.running_var.uniform_(0.5, 1.5)
    sd = {f"n.{k}": v for k, v in node.state_dict().items()}
    pl = P.Plan()
    for name, c, s in (("a", 64, 8), ("b", 128, 8), ("c", 64, 4)):
        pl.add_buffer(name, c, s)
    out = P._Lowering(pl, sd).fuse("n", "n", [("a", 64), ("b", 128), ("c", 64)], 64, 8, "down")
    fuse = next(op for op in pl.ops if op.kind == "fuse")
    w = torch.relu(node.weights.detach().double())
    assert fuse.resize == 2 and len(fuse.srcs) == 3
    assert abs(fuse.scales[0] - float(w[0] / (w.sum() + 1e-6))) < 1e-12 and fuse.scales[1] == 1.0       # b is projected: weight (0) folded
    a, b, c = torch.rand(1, 64, 8, 8), torch.rand(1, 128, 8, 8), torch.rand(1, 64, 16, 16)
    # run the op list by hand (the emulator starts from the image): same arithmetic as tests/plan_emulator.py
    import torch.nn.functional as F
    bufs = {"a": a, "b": b, "c": c}
    for op in pl.ops:
        if op.kind == "fuse":
            terms = [bufs[s] * sc for s, sc in zip(op.srcs[:-1], op.scales[:-1])] + [F.max_pool2d(bufs[op.srcs[-1]], 2, 2) * op.scales[-1]]
            bufs[op.dst] = sum(terms)
        elif op.kind == "dw":
            bufs[op.dst] = F.relu6(F.conv2d(bufs[op.src], op.weight, op.bias, 1, 1, groups=op.cin))
        else:
            y = F.conv2d(bufs[op.src], op.weight, op.bias, 1, op.pad)
            bufs[op.dst] = F.relu6(y) if op.relu == 2 else y
    with torch.no_grad():
        ref = node(a, b, c)
    np.testing.assert_allclose(bufs[out].numpy(), ref.numpy(), rtol=0, atol=2e-5)
