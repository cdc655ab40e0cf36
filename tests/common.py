"""Shared test helpers (configs, golden loading, error metrics)."""
import os

import torch

from magicdrive_b200 import arch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def tiny_configs():
    u = arch.UNetConfig(block_out_channels=(64, 128), down_block_types=("CrossAttnDownBlock2D", "DownBlock2D"),
                        up_block_types=("UpBlock2D", "CrossAttnUpBlock2D"), layers_per_block=1, attention_head_dim=2)
    c = arch.ControlNetConfig(block_out_channels=(64, 128), down_block_types=("CrossAttnDownBlock2D", "DownBlock2D"),
                              layers_per_block=1, attention_head_dim=2, map_size=(8, 52, 52))
    return u, c


def golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


def tiny_state_dicts(seed=7):
    u, c = tiny_configs()
    return (arch.synthetic_state_dict(arch.unet_param_shapes(u), seed),
            arch.synthetic_state_dict(arch.controlnet_param_shapes(c), seed + 1))


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def max_rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def to_dev(x, dev, dtype=None):
    if isinstance(x, dict):
        return {k: to_dev(v, dev, dtype) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [to_dev(v, dev, dtype) for v in x]
    if torch.is_tensor(x):
        if dtype is not None and x.is_floating_point():
            return x.to(dev, dtype)
        return x.to(dev)
    return x


def record(line: str, name: str = "parity_gpu_latest.txt"):
    """Print a `[parity]` / `[speed]` evidence line and append it to profiles/<name> (and gpurun_out/<name> when that
    directory exists), so that a GPU run of the test-suite leaves its measured numbers behind (pytest -q hides stdout)."""
    print(line)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for d in ("profiles", "gpurun_out"):
        dd = os.path.join(root, d)
        if os.path.isdir(dd):
            try:
                with open(os.path.join(dd, name), "a") as fh:
                    fh.write(line.rstrip("\n") + "\n")
            except OSError:
                pass
