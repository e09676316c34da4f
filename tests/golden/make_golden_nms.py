"""Generate tests/golden/nms.npz by running the REFERENCE'S OWN lib/utils/nms.py (plain numpy, imports under
Python 3) in this container on tests/golden/cases.nms_inputs(), followed by a literal execution of the pose
assembly loop of lib/fcn/test.py:197-211 (that file is Python 2 and cannot be imported; the six lines are
restated below, citing the lines).

    python tests/golden/make_golden_nms.py          # needs /root/reference; the .npz is committed
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.golden import cases  # noqa: E402


def main():
    spec = importlib.util.spec_from_file_location("ref_nms", "/root/reference/lib/utils/nms.py")
    ref_nms = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_nms)
    out = {}
    for tag, kw in (("a", dict(seed=77, n=96)), ("b", dict(seed=78, n=128)), ("c", dict(seed=79, n=7)), ("d", dict(seed=80, n=1))):
        rois, poses_init, poses_pred = cases.nms_inputs(**kw)
        keep = ref_nms.nms(rois, 0.5)                                   # test.py:198
        r = rois[keep, :]                                               # test.py:199
        poses = poses_init[keep, :].copy()                              # test.py:200, 207
        pp = poses_pred[keep, :]                                        # test.py:201
        for i in range(r.shape[0]):                                     # test.py:208-211
            class_id = int(r[i, 1])
            if class_id >= 0:
                poses[i, :4] = pp[i, 4 * class_id:4 * class_id + 4]
        out[f"{tag}_keep"] = np.asarray(keep, np.int32)
        out[f"{tag}_rois"] = r
        out[f"{tag}_poses"] = poses
        print(tag, "n", rois.shape[0], "kept", len(keep))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "nms.npz"), **out)


if __name__ == "__main__":
    main()
