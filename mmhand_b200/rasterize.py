"""Keypoint -> 21-joint Gaussian heatmap rasteriser on the GPU (mmh_heatmap_rasterize).

Mirrors ``Genericdataset.get_heatmaps / gen_heatmap / gaussian_kernel`` of the reference
(data/generic_dataset.py:191-199, :208-217, :238-242): one map per joint,
``exp(-((gx-x)^2+(gy-y)^2)/2/sigma/sigma)`` in float64, values > 1 clamped to 1, values < 0.0099 zeroed, cast to
float32 last. The reference passes ``(shape[0], shape[1])`` as ``(width, height)`` (harmless on square frames, Q15);
here ``shape = (H, W)`` and x indexes columns.
"""
import torch

from . import runtime

SIGMA = 6.0
THRESH = 0.0099


def get_heatmaps(uv_coords, shape=(256, 256), sigma=SIGMA, thresh=THRESH, out=None, device=None):
    """uv_coords: [..., J, 2] (x, y), any float dtype / device -> float32 [..., J, H, W] on the GPU."""
    H, W = int(shape[0]), int(shape[1])
    uv = torch.as_tensor(uv_coords)
    ops = runtime.get_ops(device if device is not None else (uv.device if uv.is_cuda else None))
    uv = uv.to(ops.device, torch.float64, non_blocking=True).contiguous()
    lead = tuple(uv.shape[:-1])
    if out is None:
        out = torch.empty(lead + (H, W), dtype=torch.float32, device=ops.device)
    ops.heatmaps(uv, H, W, sigma, thresh, out)
    return out


MISSING_VALUE = -1


def cords_to_map(cords, img_size, sigma=6, device=None):
    """``cords_to_map`` of the reference's dataset tool (tool/generate_pose_map_RHD.py:22-29) on the GPU: ``cords``
    [..., J, 2] as (y, x) -> float32 [..., H, W, J] un-thresholded Gaussian maps (HWC), a joint with a MISSING_VALUE
    coordinate keeps a zero plane. Batched: any leading dimensions."""
    H, W = int(img_size[0]), int(img_size[1])
    yx = torch.as_tensor(cords)
    ops = runtime.get_ops(device if device is not None else (yx.device if yx.is_cuda else None))
    yx = yx.to(ops.device, torch.float64).contiguous()
    lead, J = tuple(yx.shape[:-2]), yx.shape[-2]
    out = torch.empty(lead + (H, W, J), dtype=torch.float32, device=ops.device)
    ops.pose_maps(yx.reshape(-1, J, 2), H, W, sigma, MISSING_VALUE, out)
    return out


def generate_jointsmap(uv_coord, depth, width, height, channel=3, dtype=torch.float64, device=None):
    """``generate_jointsmap`` of the reference (data/generic_dataset.py:30-78) on the GPU (mmh_jointsmap_rasterize).

    uv_coord: [..., 21, 2] (x, y), depth: [..., 21] -> the part map. ``dtype=torch.float64`` returns the reference's
    canvas ``[..., height, width, 3]`` float64; ``dtype=torch.uint8`` the compact ``[..., height, width]`` map (the
    three channels are equal). Same argument order as the reference (width before height)."""
    assert channel == 3, "the reference's canvas has three equal channels"
    uv = torch.as_tensor(uv_coord)
    ops = runtime.get_ops(device if device is not None else (uv.device if uv.is_cuda else None))
    uv = uv.to(ops.device, torch.float64).contiguous()
    z = torch.as_tensor(depth).to(ops.device, torch.float64).contiguous()
    lead = tuple(uv.shape[:-2])
    assert uv.shape[-2:] == (21, 2) and tuple(z.shape) == lead + (21,)
    H, W = int(height), int(width)
    if dtype == torch.uint8:
        out = torch.empty(lead + (H, W), dtype=torch.uint8, device=ops.device)
        ops.jointsmap(uv, z, H, W, None, out)
    else:
        out = torch.empty(lead + (H, W, 3), dtype=torch.float64, device=ops.device)
        ops.jointsmap(uv, z, H, W, out, None)
    return out
