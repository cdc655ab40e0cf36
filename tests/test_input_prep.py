"""Input preparation: the oracle restatement and the CUDA kernels against the outputs of the reference's own
code on its six real nuScenes samples (tests/golden/input_prep.pt <- oracle/make_golden_input_prep.py)."""
import pytest
import torch

from oracle import input_prep as OP
from tests.common import golden


def test_oracle_box_preprocessing_matches_reference_on_its_demo_samples():
    for c in golden("input_prep.pt"):
        out = OP.preprocess_bbox(c["gt_bboxes_3d"], c["gt_labels_3d"], c["lidar2camera"], c["img_aug_matrix"])
        assert (out is None) == (c["bboxes"] is None)
        assert out["bboxes"].shape == c["bboxes"].shape
        assert torch.equal(out["masks"], c["masks"]) and torch.equal(out["classes"], c["classes"])
        assert torch.equal(out["bboxes"], c["bboxes"])  # same fp32 operations in the same order: bit-exact
        assert torch.equal(OP.camera_param(c["camera_intrinsics"], c["lidar2camera"]), c["camera_param"])


def test_oracle_fixed_capacity_pads_like_bbox_max_length():
    c = golden("input_prep.pt")[0]
    out = OP.preprocess_bbox(c["gt_bboxes_3d"], c["gt_labels_3d"], c["lidar2camera"], c["img_aug_matrix"], max_len=40)
    n = c["bboxes"].shape[1]
    assert out["bboxes"].shape == (6, 40, 8, 3) and torch.equal(out["bboxes"][:, :n], c["bboxes"])
    assert not out["masks"][:, n:].any() and (out["classes"][:, n:] == -1).all() and (out["bboxes"][:, n:] == 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("max_len", [None, 48])
def test_device_collate_matches_reference_on_its_demo_samples(cuda_lib, max_len):
    from magicdrive_b200.input_prep import collate_on_device
    cases = golden("input_prep.pt")
    examples = [dict(gt_bboxes_3d=c["gt_bboxes_3d"], gt_labels_3d=c["gt_labels_3d"], lidar2camera=c["lidar2camera"],
                     img_aug_matrix=c["img_aug_matrix"], camera_intrinsics=c["camera_intrinsics"],
                     gt_masks_bev=torch.zeros(8, 20, 20)) for c in cases]
    # one sample at a time (the reference's demo path) and the whole batch (collate_fn: padded to the batch's longest list)
    for ex, c in zip(examples, cases):
        r = collate_on_device([ex], "cuda", max_len=max_len)
        bx = r["kwargs"]["bboxes_3d_data"]
        n = c["bboxes"].shape[1]
        assert torch.allclose(r["camera_param"][0].cpu(), c["camera_param"], rtol=1e-6, atol=1e-6)
        assert bx["bboxes"].shape[2] == (n if max_len is None else max_len)
        assert torch.equal(bx["masks"][0, :, :n].cpu(), c["masks"]) and torch.equal(bx["classes"][0, :, :n].cpu(), c["classes"])
        assert torch.allclose(bx["bboxes"][0, :, :n].cpu(), c["bboxes"], rtol=1e-5, atol=2e-5)
        assert not bx["masks"][0, :, n:].any() and (bx["bboxes"][0, :, n:] == 0).all()
        assert bx["counts"][0].cpu().tolist() == c["masks"].sum(-1).tolist()
    r = collate_on_device(examples, "cuda", max_len=max_len)
    bx = r["kwargs"]["bboxes_3d_data"]
    longest = max(c["bboxes"].shape[1] for c in cases)
    assert bx["bboxes"].shape == (len(cases), 6, longest if max_len is None else max_len, 8, 3)
    for i, c in enumerate(cases):
        n = c["bboxes"].shape[1]
        assert torch.equal(bx["masks"][i, :, :n].cpu(), c["masks"]) and not bx["masks"][i, :, n:].any()
        assert torch.allclose(bx["bboxes"][i, :, :n].cpu(), c["bboxes"], rtol=1e-5, atol=2e-5)
