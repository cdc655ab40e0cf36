"""TEST INFRASTRUCTURE ONLY — golden vectors for the step BEFORE the denoising path: the reference's own
box pre-processing and camera-parameter assembly, run on the six real nuScenes samples it ships (demo/data/*.pth).

The functions are the reference's: `demo/helper.py` restates `magicdrive/dataset/utils.py:_preprocess_bbox / collate_fn`
without mmdet3d (its own copy of LiDARInstance3DBoxes etc.), but imports cv2 / omegaconf at module level, which this image does
not have.  So the needed definitions are taken from its source text with `ast` and executed unmodified.

    python -m oracle.make_golden_input_prep      # build container only (needs /root/reference)
"""
import ast
import copy
import glob
import os
import sys

import numpy as np
import torch

REF = os.environ.get("MAGICDRIVE_REFERENCE_SRC", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "input_prep.pt")
NEEDED = ["rotation_3d_in_axis", "LiDARInstance3DBoxes", "box_center_shift", "trans_boxes_to_views", "trans_boxes_to_view",
          "ensure_positive_z", "_preprocess_bbox", "precompute_cam_ext"]


def reference_namespace():
    src = open(os.path.join(REF, "demo", "helper.py")).read()
    tree = ast.parse(src)
    ns = {"np": np, "torch": torch, "copy": copy, "List": list, "Optional": object, "Tuple": tuple}
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in NEEDED:
            exec(compile(ast.Module([node], []), "demo/helper.py", "exec"), ns)
    assert all(n in ns for n in NEEDED)
    return ns


def main():
    ns = reference_namespace()
    cases = []
    for f in sorted(glob.glob(os.path.join(REF, "demo", "data", "*.pth"))):
        ex = torch.load(f, weights_only=False)
        ex = ns["precompute_cam_ext"](ex)  # camera2lidar, lidar2image (demo/helper.py:495-504)
        boxes = ns["_preprocess_bbox"](ex)  # demo/helper.py:386-466 == dataset/utils.py:120-240 for one sample, 3-D filter
        cam = torch.cat([ex["camera_intrinsics"][:, :3, :3], ex["camera2lidar"][:, :3]], dim=-1)  # :544-547 == utils.py:294-297
        cases.append(dict(name=os.path.basename(f), gt_bboxes_3d=ex["gt_bboxes_3d"].clone(), gt_labels_3d=ex["gt_labels_3d"].clone(),
                          camera_intrinsics=ex["camera_intrinsics"].clone(), lidar2camera=ex["lidar2camera"].clone(),
                          img_aug_matrix=ex["img_aug_matrix"].clone(), camera_param=cam.clone(),
                          bboxes=None if boxes is None else boxes["bboxes"][0].clone(),
                          classes=None if boxes is None else boxes["classes"][0].clone(),
                          masks=None if boxes is None else boxes["masks"][0].clone()))
        print(cases[-1]["name"], "boxes", tuple(ex["gt_bboxes_3d"].shape), "->", None if boxes is None else tuple(boxes["bboxes"].shape),
              "visible per view", None if boxes is None else boxes["masks"][0].sum(-1).tolist())
    torch.save(cases, OUT)
    print(OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    sys.exit(main())
