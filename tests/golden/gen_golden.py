"""Generate the committed golden vectors by running the UNMODIFIED reference in the authoring container.

    python tests/golden/gen_golden.py        (needs /root/reference; the GPU box only reads the .npz files)

Decode goldens: outputs of the reference's own ``CenterNet.decode_detections`` /
``get_topk_from_heatmap`` / ``gather_and_decode_boxes`` (reference models/centernet.py:229-304),
imported under stub modules by oracle/ref_import.py.  Inputs are NOT stored: they are regenerated
from the seeded recipe in tests/cases.py, so each file holds only the small outputs.

Forward golden: the reference's own ``GenericModel`` + ``GenericHead`` (reference models/meta.py:21-47)
wrapped around the in-repo backbone/neck stand-ins (vision_toolbox is not vendored), on a
64x64 seeded image - pins the wiring of the spec model and gives the GPU engine a value-level target.

Tracker goldens: the reference's own ``Tracker`` (reference models/tracker.py, loaded by oracle/tracker_np.py) driven
through ``update`` with seeded detection streams; every live track of every frame is recorded.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_import, spec_model, tracker_np  # noqa: E402
import cases  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    assert ref_import.reference_available(), "run this where /root/reference exists"
    torch.set_num_threads(4)
    for case in cases.DECODE_CASES:
        if len(sys.argv) > 1:
            break
        heat, box, reid = cases.make_decode_inputs(case)
        probs = heat.sigmoid() if case["logits"] else heat
        out = ref_import.reference_decode(probs, box, num_detections=case["k"], nms_kernel=case["nms"],
                                          normalize_boxes=case["normalize"], box_log=case["box_log"],
                                          box_multiplier=case["mult"], stride=case["stride"])
        arrs = {k: v.numpy() for k, v in out.items()}
        np.savez_compressed(os.path.join(HERE, f"decode_{case['name']}.npz"), **arrs)
        print("decode", case["name"], {k: v.shape for k, v in arrs.items()})

    only = set(sys.argv[1:])            # python gen_golden.py forward_<name> ...: (re)generate just those files
    # IDA / BiFPN necks: the fusion nodes and separable convs of the golden model are the reference's OWN Fuse / make_conv
    # classes (models/layers.py:40-79, 138-177) - the oracle's restatements are swapped out while the golden is made
    ref_layers = ref_import.import_reference_layers()
    for name, kw in cases.FORWARD_CASES.items():
        if only and f"forward_{name}" not in only:
            continue
        restated = spec_model.Fuse, spec_model.make_conv
        spec_model.Fuse, spec_model.make_conv = ref_layers.Fuse, ref_layers.make_conv
        try:
            m = spec_model.synth_init(spec_model.build_spec_model(**kw["model"]), seed=kw["seed"], **kw.get("init", {}))
        finally:
            spec_model.Fuse, spec_model.make_conv = restated
        rm = ref_import.reference_generic_model(m)
        x = cases.make_image(kw)
        with torch.no_grad():
            out = rm(x)
        np.savez_compressed(os.path.join(HERE, f"forward_{name}.npz"), **{k: v.numpy() for k, v in out.items()})
        print("forward", name, {k: tuple(v.shape) for k, v in out.items()})

    # tracker goldens: the reference's own Tracker.update (models/tracker.py:131-201, scipy cdist + utils/box.py costs,
    # scipy Hungarian) on the seeded detection streams of tests/cases.py; rows = (frame, track_id, state, bbox)
    ref_trk = tracker_np.import_reference_tracker()
    import warnings
    for name, case in cases.TRACK_CASES.items():
        if len(sys.argv) > 1:
            break
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t = ref_trk.Tracker(model=None, **case["tracker"])
        rows = cases.run_track_sequence(t, cases.make_track_sequence(case))
        np.savez_compressed(os.path.join(HERE, f"tracker_{name}.npz"), rows=rows, next_track_id=np.array(t.next_track_id))
        print("tracker", name, rows.shape, "tracks created:", t.next_track_id)


if __name__ == "__main__":
    main()
