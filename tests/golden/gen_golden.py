"""Generate the committed golden vectors by running the UNMODIFIED reference in the authoring container.

    python tests/golden/gen_golden.py        (needs /root/reference; the GPU box only reads the .npz files)

Decode goldens: outputs of the reference's own ``CenterNet.decode_detections`` /
``get_topk_from_heatmap`` / ``gather_and_decode_boxes`` (reference models/centernet.py:229-304),
imported under stub modules by oracle/ref_import.py.  Inputs are NOT stored: they are regenerated
from the seeded recipe in tests/cases.py, so each file holds only the small outputs.

Forward golden: the reference's own ``GenericModel`` + ``GenericHead`` (reference models/meta.py:21-47)
wrapped around the in-repo backbone/neck stand-ins (vision_toolbox is not vendored), on a
64x64 seeded image - pins the wiring of the spec model and gives the GPU engine a value-level target.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_import, spec_model  # noqa: E402
import cases  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    assert ref_import.reference_available(), "run this where /root/reference exists"
    torch.set_num_threads(4)
    for case in cases.DECODE_CASES:
        heat, box, reid = cases.make_decode_inputs(case)
        probs = heat.sigmoid() if case["logits"] else heat
        out = ref_import.reference_decode(probs, box, num_detections=case["k"], nms_kernel=case["nms"],
                                          normalize_boxes=case["normalize"], box_log=case["box_log"],
                                          box_multiplier=case["mult"], stride=case["stride"])
        arrs = {k: v.numpy() for k, v in out.items()}
        np.savez_compressed(os.path.join(HERE, f"decode_{case['name']}.npz"), **arrs)
        print("decode", case["name"], {k: v.shape for k, v in arrs.items()})

    for name, kw in cases.FORWARD_CASES.items():
        m = spec_model.synth_init(spec_model.build_spec_model(**kw["model"]), seed=kw["seed"])
        rm = ref_import.reference_generic_model(m)
        x = cases.make_image(kw)
        with torch.no_grad():
            out = rm(x)
        np.savez_compressed(os.path.join(HERE, f"forward_{name}.npz"), **{k: v.numpy() for k, v in out.items()})
        print("forward", name, {k: tuple(v.shape) for k, v in out.items()})


if __name__ == "__main__":
    main()
