"""Seeded synthetic inputs of SURVEY.md §8(d) (latents, CLIP-like text states, camera K | cam2lidar, 3D boxes,
BEV map).  Pure data generation shared by bench.py, tests and oracle/make_golden.py: no model arithmetic here."""
import torch


def synthetic_inputs(scenes, n_cam, h, w, n_box, map_hw, seed=0, text_len=77):
    g = torch.Generator().manual_seed(seed)
    lat = torch.randn(scenes, 4, h, w, generator=g)
    text = torch.randn(scenes, text_len, 768, generator=g)
    neg = torch.randn(1, text_len, 768, generator=g).expand(scenes, -1, -1).contiguous()
    K = torch.tensor([[1266.0, 0, 816.0], [0, 1266.0, 491.0], [0, 0, 1.0]])
    cams = []
    for yaw_deg in [55.0, 0.0, -55.0, -110.0, 180.0, 110.0][:n_cam]:
        y = torch.deg2rad(torch.tensor(yaw_deg))
        R = torch.tensor([[torch.cos(y), -torch.sin(y), 0.0], [torch.sin(y), torch.cos(y), 0.0], [0.0, 0.0, 1.0]])
        t = (torch.rand(3, 1, generator=g) - 0.5) * 3.0
        cams.append(torch.cat([K, R, t], dim=1))  # (3, 7)
    cam = torch.stack(cams)[None].expand(scenes, -1, -1, -1).contiguous()
    boxes = None
    if n_box > 0:
        centre = torch.cat([(torch.rand(scenes, n_cam, n_box, 1, 2, generator=g) - 0.5) * 100,
                            (torch.rand(scenes, n_cam, n_box, 1, 1, generator=g) - 0.5) * 4], -1)
        size = torch.rand(scenes, n_cam, n_box, 1, 3, generator=g) * 4.5 + 0.5
        sign = torch.tensor([[sx, sy, sz] for sx in (-.5, .5) for sy in (-.5, .5) for sz in (-.5, .5)])
        corners = centre + size * sign[None, None, None]
        nvalid = torch.randint(max(1, n_box // 4), n_box + 1, (scenes, n_cam, 1), generator=g)
        masks = torch.arange(n_box)[None, None] < nvalid
        boxes = {"bboxes": corners.contiguous(), "classes": torch.randint(0, 10, (scenes, n_cam, n_box), generator=g),
                 "masks": masks}
    bev = (torch.rand(scenes, 8, map_hw, map_hw, generator=g) < 0.15).float()
    return dict(latents=lat, prompt_embeds=text, negative_prompt_embeds=neg, camera_param=cam, bboxes_3d_data=boxes,
                bev_map=bev)
