"""Single-op probes of the tcgen05 conv kernel against torch (fp64 on CPU).  Each probe runs in its own process
(`python tools/conv_probe.py <index>`), so a hang or fault in one configuration does not hide the others:

    for i in $(seq 0 $(python tools/conv_probe.py count)); do timeout 90 python tools/conv_probe.py $i; done
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centernet_lightning_b200.plan import Plan, ConvOp  # noqa: E402


def cfg(name, cin, cout, k, stride, hw, n=2, relu=True, res=0, precision=0, nchw=False, src_c=None, src_off=0, dst_c=None,
        dst_off=0, up=0, deconv=0):
    """up=1: conv + nearest x2 (upsampling store); deconv=k: ConvTranspose2d(k, stride 2) as four sub-pixel phase convs."""
    return dict(name=name, cin=cin, cout=cout, k=k, stride=stride, hw=hw, n=n, relu=relu, res=res, precision=precision,
                nchw=nchw, src_c=src_c or cin, src_off=src_off, dst_c=dst_c or cout, dst_off=dst_off, up=up, deconv=deconv)


PROBES = [
    cfg("1x1_64_64_w128_fast", 64, 64, 1, 1, 128, precision=1, relu=False),
    cfg("1x1_64_64_w128_split", 64, 64, 1, 1, 128, relu=False),
    cfg("3x3_64_64_w128_fast", 64, 64, 3, 1, 128, precision=1),
    cfg("3x3_64_64_w128_split", 64, 64, 3, 1, 128),
    cfg("3x3_256_256_w128_split", 256, 256, 3, 1, 128, n=1),
    cfg("3x3_256_256_w32_split", 256, 256, 3, 1, 32),
    cfg("3x3_128_128_w64_split", 128, 128, 3, 1, 64),
    cfg("3x3_512_512_w16_split", 512, 512, 3, 1, 16),
    cfg("3x3_512_512_w8_split", 512, 512, 3, 1, 8),
    cfg("1x1_512_256_w16_split", 512, 256, 1, 1, 16, relu=False),
    cfg("3x3s2_64_128_w64_split", 64, 128, 3, 2, 64),
    cfg("1x1s2_64_128_w64_split", 64, 128, 1, 2, 64, relu=False),
    cfg("3x3s2_256_512_w16_split", 256, 512, 3, 2, 16),
    cfg("3x3_64_64_res_split", 64, 64, 3, 1, 128, res=1),
    cfg("1x1_64_256_resup2_split", 64, 256, 1, 1, 128, relu=False, res=2),
    cfg("1x1_256_80_nchw_split", 256, 80, 1, 1, 128, relu=False, nchw=True),
    cfg("1x1_256_4_nchw_split", 256, 4, 1, 1, 128, relu=False, nchw=True),
    cfg("3x3_256_512_w128_split", 256, 512, 3, 1, 128, n=1),
    cfg("3x3_group_off256_split", 256, 256, 3, 1, 64, src_c=512, src_off=256, dst_c=512, dst_off=256),
    cfg("3x3_256_256_w256_split", 256, 256, 3, 1, 256, n=1),
    cfg("3x3_64_64_w272_split", 64, 64, 3, 1, (32, 272), n=1),
    cfg("3x3_256_256_w128_fast", 256, 256, 3, 1, 128, n=1, precision=1),
    cfg("3x3_512_256_up2_w16_split", 512, 256, 3, 1, 16, up=1),
    cfg("3x3_128_64_up2_w64_split", 128, 64, 3, 1, 64, up=1),
    cfg("3x3_64_64_up2_w24x40_split", 64, 64, 3, 1, (24, 40), n=1, up=1),
    cfg("deconv3_256_w16_split", 256, 256, 3, 1, 16, deconv=3),
    cfg("deconv4_64_w64_split", 64, 64, 4, 1, 64, deconv=4),
    cfg("deconv4_128_w24x40_split", 128, 128, 4, 1, (24, 40), n=1, deconv=4),
    # row-rolling form (Cin = 64, rows wider than 64 pixels): ragged widths, short maps, batch crossing CTA runs,
    # non-square phase windows and the upsampling store
    cfg("rows_3x3_64_64_w104x96_split", 64, 64, 3, 1, (96, 104), n=3),
    cfg("rows_3x3_64_128_h8_w128_split", 64, 128, 3, 1, (8, 128), n=5),
    cfg("rows_3x3_64_64_h8_w72_split", 64, 64, 3, 1, (8, 72), n=2, relu=False),
    cfg("rows_3x3_64_64_w200_res_split", 64, 64, 3, 1, (40, 200), n=2, res=1),
    cfg("rows_deconv4_64_w128_split", 64, 64, 4, 1, (24, 128), n=2, deconv=4),
    cfg("rows_deconv3_64_w72_split", 64, 64, 3, 1, (16, 72), n=1, deconv=3),
    cfg("rows_3x3_64_64_up2_w128_split", 64, 64, 3, 1, (16, 128), n=2, up=1),
    # CTA-pair form (run with CNL_PAIR_MIN_TILES=2 so that these small problems take it): odd tile counts leave the
    # peer CTA of the last cluster with a dummy tile; two Cout tiles; residual; stride 2
    cfg("pair_3x3_256_256_odd9tiles_split", 256, 256, 3, 1, (40, 8), n=3),
    cfg("pair_3x3_256_512_odd3tiles_res_split", 256, 512, 3, 1, (40, 8), n=1, res=1),
    cfg("pair_3x3s2_256_256_w64_split", 256, 256, 3, 2, 64, n=1),
    cfg("pair_3x3_512_256_w32_1img_split", 512, 256, 3, 1, 32, n=1),
]


def run(c, device="cuda:0", verbose=True):
    from centernet_lightning_b200.engine import Engine
    g = torch.Generator().manual_seed(hash(c["name"]) % 1000)
    hw = c["hw"] if isinstance(c["hw"], tuple) else (c["hw"], c["hw"])
    oh, ow = hw
    ih, iw = oh * c["stride"], ow * c["stride"]
    # engine geometry: image size = input map size x 1 (buffer stride 1) rounded so that everything divides
    f = 4                                   # the engine wants an image that is a multiple of 32: maps live at stride >= 4
    H, W = ih * f, iw * f
    p = Plan()
    p.add_buffer("image", 3, 1, fp32_nchw=True)
    p.add_buffer("src", c["src_c"], f)
    p.add_buffer("dst", c["dst_c"], f * c["stride"], fp32_nchw=c["nchw"])
    res_name = None
    if c["res"]:
        p.add_buffer("res", c["cout"], f * c["stride"] * c["res"])
        res_name = "res"
    k = c["k"]
    up = 2 if (c["up"] or c["deconv"]) else 1
    if up == 2:
        p.buffers["dst"].stride = f // 2                # destination at twice the conv's resolution
    b = torch.randn((c["cout"],), generator=g) * 0.1
    if c["deconv"]:
        from centernet_lightning_b200.plan import lower_conv_transpose
        w = torch.randn((c["cin"], c["cout"], k, k), generator=g) * (8.0 / (c["cin"] * k * k)) ** 0.5
        p.ops = lower_conv_transpose("probe", "src", "dst", w, None, relu=c["relu"])
        for op in p.ops:
            op.bias = b
    else:
        w = torch.randn((c["cout"], c["cin"], k, k), generator=g) * (2.0 / (c["cin"] * k * k)) ** 0.5
        p.ops.append(ConvOp("probe", "conv", "src", "dst", c["cin"], c["cout"], k, c["stride"], k // 2, w, b, relu=c["relu"],
                            src_c_off=c["src_off"], dst_c_off=c["dst_off"], residual=res_name, residual_up=max(1, c["res"]),
                            dst_up=up, dst_phase=-1))
    p.outputs = {}
    dev = torch.device(device)
    eng = Engine(p, c["n"], H, W, dev, precision=c["precision"])
    x = torch.randn((c["n"], c["src_c"], ih, iw), generator=g)
    eng.write_buffer("src", x)
    r = None
    if c["res"]:
        r = torch.randn((c["n"], c["cout"], oh // c["res"], ow // c["res"]), generator=g)
        eng.write_buffer("res", r)
    if not c["nchw"]:
        eng.write_buffer("dst", torch.full((c["n"], c["dst_c"], oh * up, ow * up), 7.0))      # sentinel: untouched channels must survive
    eng.forward(None)
    torch.cuda.synchronize()
    got = eng.read_buffer("dst").cpu()
    # reference: what the kernel sees is the fp16 (hi[+lo]) rounding of x / w
    def q(t):
        hi = t.half().float()
        return hi if c["precision"] == 1 else hi + (t - hi).half().float()
    xin = q(x)[:, c["src_off"]:c["src_off"] + c["cin"]].double()
    if c["deconv"]:
        op_pad = k % 2
        ref = F.conv_transpose2d(xin, w.double(), None, stride=2, padding=(k + op_pad) // 2 - 1, output_padding=op_pad)
        ref = ref + b.double().view(1, -1, 1, 1)
    else:
        ref = F.conv2d(xin, w.double(), None, c["stride"], k // 2) + b.double().view(1, -1, 1, 1)
    if c["up"]:
        ref = F.interpolate(ref, scale_factor=2.0, mode="nearest")
    if r is not None:
        rr = q(r).double()
        if c["res"] == 2:
            rr = F.interpolate(rr, scale_factor=2.0, mode="nearest")
        ref = ref + rr
    if c["relu"]:
        ref = ref.clamp_min(0)
    out = got[:, c["dst_off"]:c["dst_off"] + c["cout"]].double()
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    untouched_ok = True
    if not c["nchw"] and c["dst_c"] != c["cout"]:
        mask = torch.ones(c["dst_c"], dtype=torch.bool)
        mask[c["dst_off"]:c["dst_off"] + c["cout"]] = False
        untouched_ok = bool((got[:, mask] == 7.0).all())
    tol = (2e-2 if c["precision"] == 1 else 2e-4) * max(1.0, scale)
    ok = err < tol and untouched_ok
    if verbose:
        print(f"{'OK ' if ok else 'BAD'} {c['name']:32s} max_err={err:.3e} ref_max={scale:.2f} tol={tol:.1e} untouched={untouched_ok}", flush=True)
        if not ok:
            d = (out - ref).abs()
            nz = (d > tol).nonzero()
            print("   first bad idx:", nz[:5].tolist(), "n_bad", len(nz), "of", d.numel(), flush=True)
            print("   got", out.flatten()[:8].tolist(), "\n   ref", ref.flatten()[:8].tolist(), flush=True)
    eng.close()
    return ok, err


if __name__ == "__main__":
    if sys.argv[1] == "count":
        print(len(PROBES) - 1)
    else:
        run(PROBES[int(sys.argv[1])])
