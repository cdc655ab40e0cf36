"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's input preparation for one sample (the step before the
denoising path).  Pinned to the reference's own code on its six real samples by
tests/golden/input_prep.pt (oracle/make_golden_input_prep.py).  Nothing in the product imports this file.

Follows, line by line:
  corners()            mmdet3d LiDARInstance3DBoxes.corners as vendored in demo/helper.py:152-190, z-rotation :39-85
  preprocess_bbox()    magicdrive/dataset/utils.py:120-240 (bbox_mode 'all-xyz', view_shared False, use_3d_filter True)
                       == demo/helper.py:386-466; the visibility test projects the box whose origin was re-interpreted as the
                       gravity centre (box_center_shift, demo/helper.py:310-314 / runner/utils.py) with
                       trans = img_aug_matrix @ lidar2camera in float64, ensure_positive_z (dataset/utils.py:69-71)
  camera_param()       dataset/utils.py:294-297 with camera2lidar from demo/helper.py:495-501
"""
import numpy as np
import torch


def corners(boxes: torch.Tensor) -> torch.Tensor:
    """(n, >=7) rows (x, y, z, dx, dy, dz, yaw, ...) bottom-centred -> (n, 8, 3) fp32."""
    dims = boxes[:, 3:6]
    norm = torch.from_numpy(np.stack(np.unravel_index(np.arange(8), [2] * 3), axis=1)).to(dims.dtype)
    norm = norm[[0, 1, 3, 2, 4, 5, 7, 6]] - dims.new_tensor([0.5, 0.5, 0])
    c = dims.view(-1, 1, 3) * norm.reshape(1, 8, 3)
    sin, cos = torch.sin(boxes[:, 6]), torch.cos(boxes[:, 6])
    one, zero = torch.ones_like(cos), torch.zeros_like(cos)
    rot_t = torch.stack([torch.stack([cos, -sin, zero]), torch.stack([sin, cos, zero]), torch.stack([zero, zero, one])])
    c = torch.einsum("aij,jka->aik", c, rot_t)
    return c + boxes[:, :3].view(-1, 1, 3)


def preprocess_bbox(gt_bboxes_3d: torch.Tensor, gt_labels_3d: torch.Tensor, lidar2camera: torch.Tensor,
                    img_aug_matrix: torch.Tensor, max_len=None):
    """-> dict(bboxes (V, L, 8, 3), classes (V, L) int64 (-1 padding), masks (V, L) bool) or None; L = max visible count
    over the views unless `max_len` fixes it."""
    if len(gt_bboxes_3d) == 0:
        return None
    shifted = gt_bboxes_3d.clone().float()
    shifted[:, :3] += shifted[:, 3:6] * (shifted.new_tensor((0.5, 0.5, 0)) - shifted.new_tensor((0.5, 0.5, 0.5)))
    cs = corners(shifted).numpy()
    homo = np.concatenate([cs.reshape(-1, 3), np.ones((cs.shape[0] * 8, 1))], axis=-1)  # float64
    index_list = []
    for v in range(lidar2camera.shape[0]):
        trans = img_aug_matrix[v].numpy() @ lidar2camera[v].numpy()  # float32
        z = (homo @ trans.reshape(4, 4).T)[:, 2].reshape(-1, 8)
        index_list.append(np.any(z > 0, axis=1))
    longest = max(int(m.sum()) for m in index_list)
    if longest == 0:
        return None
    L = longest if max_len is None else max_len
    pts = corners(gt_bboxes_3d.float())
    V = len(index_list)
    out = dict(bboxes=torch.zeros(V, L, 8, 3), classes=-torch.ones(V, L, dtype=torch.long), masks=torch.zeros(V, L, dtype=torch.bool))
    for v, m in enumerate(index_list):
        m = torch.from_numpy(m)
        n = int(m.sum())
        out["bboxes"][v, :n] = pts[m]
        out["classes"][v, :n] = gt_labels_3d[m]
        out["masks"][v, :n] = True
    return out


def camera_param(camera_intrinsics: torch.Tensor, lidar2camera: torch.Tensor) -> torch.Tensor:
    """(V, 4, 4), (V, 4, 4) -> (V, 3, 7)."""
    c2l = torch.stack([torch.eye(4, dtype=lidar2camera.dtype)] * len(lidar2camera))
    c2l[:, :3, :3] = lidar2camera[:, :3, :3].transpose(1, 2)
    c2l[:, :3, 3:] = torch.bmm(-c2l[:, :3, :3], lidar2camera[:, :3, 3:])
    return torch.cat([camera_intrinsics[:, :3, :3], c2l[:, :3]], dim=-1)
