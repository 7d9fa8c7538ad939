"""CPU: pin the oracle against the golden vectors produced by the reference's own code."""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import decode_np, decode_torch, ref_import, spec_model

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _gold(name):
    return dict(np.load(os.path.join(GOLD, f"{name}.npz")))


def _kw(case):
    return dict(num_detections=case["k"], nms_kernel=case["nms"], normalize_boxes=case["normalize"],
                box_log=case["box_log"], box_multiplier=case["mult"], stride=case["stride"])


@pytest.mark.parametrize("case", cases.DECODE_CASES, ids=lambda c: c["name"])
def test_numpy_oracle_matches_reference_golden(case):
    heat, box, reid = cases.make_decode_inputs(case)
    probs = heat.sigmoid() if case["logits"] else heat
    gold = _gold(f"decode_{case['name']}")
    out = decode_np.decode_detections(probs.numpy(), box.numpy(), **_kw(case))
    # scores bit-exact; rows identical wherever the reference's order is defined (distinct scores)
    box_tol = 2e-6 * max(1.0, float(np.abs(gold["boxes"]).max())) if case["box_log"] else 0.0   # np.exp vs ATen exp: 1 ulp
    ok, msg = decode_np.same_detections(out, gold, box_tol=box_tol)
    assert ok, msg
    if case["kind"] == "randn" and case["k"] < case["h"] * case["w"]:
        # tie-free recipe: full bit-exact agreement including order
        assert np.array_equal(out["indices"], gold["indices"])
        assert np.array_equal(out["labels"], gold["labels"])
        if not case["box_log"]:
            assert np.array_equal(out["boxes"], gold["boxes"])


@pytest.mark.parametrize("case", cases.DECODE_CASES, ids=lambda c: c["name"])
def test_aten_port_matches_reference_golden(case):
    heat, box, reid = cases.make_decode_inputs(case)
    probs = heat.sigmoid() if case["logits"] else heat
    gold = _gold(f"decode_{case['name']}")
    out = decode_torch.decode_detections(probs, box, **_kw(case))
    for k in ("scores", "indices", "labels"):
        assert np.array_equal(out[k].numpy(), gold[k]), k          # same ATen ops -> same bits, same tie order
    if case["box_log"]:     # ATen's vectorised exp differs in the last ulp between its SIMD body and scalar tail
        np.testing.assert_allclose(out["boxes"].numpy(), gold["boxes"], rtol=0, atol=4e-6 * float(np.abs(gold["boxes"]).max()))
    else:
        assert np.array_equal(out["boxes"].numpy(), gold["boxes"])


def test_embedding_gather_restatement():
    case = cases.DECODE_BY_NAME["track128"]
    heat, box, reid = cases.make_decode_inputs(case)
    out = decode_np.decode_detections(heat.sigmoid().numpy(), box.numpy(), reid=reid.numpy(), **_kw(case))
    idx = out["indices"]
    emb = out["embeddings"]
    assert emb.shape == (1, 100, 64)
    flat = reid.numpy().reshape(1, 64, -1)
    for j in (0, 17, 99):
        assert np.array_equal(emb[0, j], flat[0, :, idx[0, j]])
    t = decode_torch.gather_embeddings(reid, torch.from_numpy(idx))
    assert np.array_equal(t.numpy(), emb)


def test_tie_order_is_canonical():
    h = np.zeros((1, 2, 4, 4), np.float32)
    h[0, 0, 0, 0] = 0.5
    h[0, 1, 3, 3] = 0.5          # equal score, later index
    h[0, 1, 1, 2] = 0.9
    s, i, l = decode_np.topk_from_heatmap(h, 3, 3)
    assert i[0].tolist() == [6, 0, 15] and l[0].tolist() == [1, 0, 1]


@pytest.mark.skipif(not ref_import.reference_available(), reason="/root/reference not present (GPU box)")
def test_restatements_match_live_reference():
    g = torch.Generator().manual_seed(123)
    heat = torch.rand((2, 9, 24, 40), generator=g)
    box = torch.randn((2, 4, 24, 40), generator=g)
    ref = ref_import.reference_decode(heat, box, num_detections=30, box_multiplier=16.0)
    out = decode_np.decode_detections(heat.numpy(), box.numpy(), num_detections=30, box_multiplier=16.0)
    for k in ("scores", "indices", "labels", "boxes"):
        assert np.array_equal(ref[k].numpy(), out[k]), k


def test_spec_model_param_counts_match_published():
    """docs/experiments.md:27 of the reference: ResNet-34 21.3M, FPN(256) 2.0M, heads(256x3) 3.6M."""
    m = spec_model.build_spec_model(80)
    assert spec_model.count_params(m.backbone) == 21_284_672
    assert spec_model.count_params(m.neck) == 2_017_792
    assert spec_model.count_params(m.heads) == 3_563_604
    assert m.stride == 4
    small = spec_model.build_spec_model(80, neck_config={"out_channels": 128}, head_config={"width": 128, "depth": 2})
    assert round(spec_model.count_params(small.neck) / 1e6, 1) == 0.6          # docs/experiments.md:24
    assert round(spec_model.count_params(small.heads) / 1e6, 1) == 0.6


@pytest.mark.parametrize("name", list(cases.FORWARD_CASES))
def test_spec_model_matches_reference_generic_model_golden(name):
    kw = cases.FORWARD_CASES[name]
    m = spec_model.synth_init(spec_model.build_spec_model(**kw["model"]), seed=kw["seed"], **kw.get("init", {}))
    with torch.no_grad():
        out = m(cases.make_image(kw))
    gold = _gold(f"forward_{name}")
    assert list(out) == list(gold)
    for k in out:
        np.testing.assert_allclose(out[k].numpy(), gold[k], rtol=0, atol=1e-4)   # fp32 summation-order noise (thread count)
    assert out["heatmap"].shape[-1] == kw["size"] // 4                     # reference tests/test_models.py:68-86


def test_simple_neck_shape_contract():
    """reference tests/test_necks.py:23-38: stride-32 (4,512,16,16) -> (4,64,128,128), upsample stride 8."""
    nk = spec_model.SimpleNeck([64, 64, 128, 256, 512][1:], (256, 128, 64)).eval()
    with torch.no_grad():
        y = nk([torch.rand(4, 512, 16, 16)])
    assert tuple(y.shape) == (4, 64, 128, 128) and nk.stride == 8


def test_preprocess_oracle_properties():
    """oracle/preprocess_np.py (A.Normalize restated): float32 output, exact for the documented formula, affine and
    monotone per channel, and equal to the textbook (x/255 - mean)/std within 2 ulp."""
    from oracle import preprocess_np
    img = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, axis=2)
    out = preprocess_np.normalize(img)
    assert out.dtype == np.float32 and out.shape == (16, 16, 3)
    for c in range(3):
        col = out[..., c].reshape(-1)
        assert np.all(np.diff(col) > 0)
        textbook = (np.arange(256) / 255.0 - preprocess_np.MEAN[c]) / preprocess_np.STD[c]
        np.testing.assert_allclose(col, textbook, rtol=3e-7, atol=3e-7)
    chw = preprocess_np.to_chw(out)
    assert chw.shape == (3, 16, 16) and chw.flags["C_CONTIGUOUS"]


@pytest.mark.skipif(not ref_import.reference_available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("conv_type,weighted,resize,chans", [("normal", False, "up", [64, 128]), ("separable", True, "up", [32, 32]),
                                                             ("normal", True, "down", [64, 64, 64]), ("separable", False, "down", [24, 64])])
def test_fuse_and_make_conv_restatements_match_live_reference(conv_type, weighted, resize, chans):
    """oracle.spec_model.Fuse / make_conv against the reference's own classes (models/layers.py:40-79, 138-177) with the same
    parameters: identical outputs."""
    layers = ref_import.import_reference_layers()
    torch.manual_seed(5)
    ref = layers.Fuse(chans, 64, resize, conv_type=conv_type, weighted_fusion=weighted).eval()
    mine = spec_model.Fuse(chans, 64, resize, conv_type=conv_type, weighted_fusion=weighted).eval()
    with torch.no_grad():
        for mod in ref.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.2); mod.running_var.uniform_(0.5, 1.5); mod.weight.uniform_(0.5, 1.5); mod.bias.normal_(0, 0.1)
        if weighted:
            ref.weights.copy_(torch.rand(len(chans)) + 0.2)
    mine.load_state_dict(ref.state_dict(), strict=True)
    s = 8
    xs = [torch.randn(2, c, s, s) for c in chans[:-1]] + [torch.randn(2, chans[-1], s // 2 if resize == "up" else s * 2, s // 2 if resize == "up" else s * 2)]
    with torch.no_grad():
        assert torch.equal(ref(*xs), mine(*xs))
