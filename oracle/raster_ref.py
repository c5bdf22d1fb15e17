"""ORACLE (test infrastructure): CPU restatement of the reference's keypoint -> heatmap rasteriser.

Follows data/generic_dataset.py:191-199 (get_heatmaps: one map per joint, stacked, float32),
:208-217 (gen_heatmap: clamp >1 to 1, then zero everything < 0.0099, in float64, cast last) and
:238-242 (gaussian_kernel: mgrid over (height, width), exp(-D2 / 2.0 / sigma / sigma), same evaluation order).
Pinned: the reference's dataset module imports with three shims (oracle/ref_shims.py::load_reference_dataset_module)
and oracle/make_golden_raster.py stores outputs of its own ``Genericdataset.get_heatmaps`` for seeded and adversarial
poses in tests/golden/heatmaps_ref.npz; tests/test_raster.py requires this restatement (and the kernel body) to
reproduce them bit for bit, next to derived known answers.
"""
import numpy as np


def gaussian_kernel(width, height, x, y, sigma):
    gridy, gridx = np.mgrid[0:height, 0:width]
    D2 = (gridx - x) ** 2 + (gridy - y) ** 2
    return np.exp(-D2 / 2.0 / sigma / sigma)


def gen_heatmap(x, y, shape, sigma, thresh=0.0099):
    # the reference passes (shape[0], shape[1]) as (width, height): Q15, harmless on square frames
    m = gaussian_kernel(shape[0], shape[1], x, y, sigma)
    m[m > 1] = 1
    m[m < thresh] = 0
    return m


def get_heatmaps(uv, shape=(256, 256), sigma=6.0, thresh=0.0099):
    """uv: [J, 2] float64 (x, y) -> [J, H, W] float32."""
    return np.stack([gen_heatmap(float(x), float(y), shape, sigma, thresh).astype(np.float32) for x, y in uv])


def get_heatmaps_batch(uv, shape=(256, 256), sigma=6.0, thresh=0.0099):
    """uv: [N, J, 2] -> [N, J, H, W] float32 (vectorised, same arithmetic order)."""
    uv = np.asarray(uv, dtype=np.float64)
    H, W = shape[1], shape[0]
    gy = np.arange(H, dtype=np.float64)[None, None, :, None]
    gx = np.arange(W, dtype=np.float64)[None, None, None, :]
    D2 = (gx - uv[..., 0][..., None, None]) ** 2 + (gy - uv[..., 1][..., None, None]) ** 2
    m = np.exp(-D2 / 2.0 / sigma / sigma)
    m[m > 1] = 1
    m[m < thresh] = 0
    return m.astype(np.float32)


def normalize_image_u8(img_u8, bgr=False):
    """uint8 frames [..., H, W, 3] -> float32 [..., 3, H, W] in [-1, 1]: cv2.cvtColor(BGR2RGB) (a channel flip),
    ``normalize`` in float64 and ``make_tensor``'s ``.float()`` of data/generic_dataset.py:140-143,182-189."""
    img = np.asarray(img_u8)
    if bgr:
        img = img[..., ::-1]
    out = ((img / 255.0) - 0.5) / 0.5                                     # float64
    return np.ascontiguousarray(np.moveaxis(out, -1, -3)).astype(np.float32)


def decode_depth_u8(depth_u8, div=700.0):
    """uint8 depth frames [..., H, W, 3] as cv2.imread returns them (BGR) -> float32 [..., 3, H, W]:
    ``256.0 * px[1] + px[2]``, stacked x3, ``((d / 700.0) - 0.5) / 0.5`` in float64 (data/generic_dataset.py:148-159),
    then the fp32 copy set_input makes."""
    d = np.asarray(depth_u8)
    depth = 256.0 * d[..., 1] + d[..., 2]                                 # float64
    out = ((np.stack([depth, depth, depth], axis=-3) / div) - 0.5) / 0.5
    return out.astype(np.float32)


MISSING_VALUE = -1


def cords_to_map(cords, img_size, sigma=6):
    """tool/generate_pose_map_RHD.py:22-29, restated line by line (cords = (y, x) per joint, HWC float32 result)."""
    result = np.zeros(tuple(img_size) + cords.shape[0:1], dtype='float32')
    for i, point in enumerate(cords):
        if point[0] == MISSING_VALUE or point[1] == MISSING_VALUE:
            continue
        xx, yy = np.meshgrid(np.arange(img_size[1]), np.arange(img_size[0]))
        result[..., i] = np.exp(-((yy - point[0]) ** 2 + (xx - point[1]) ** 2) / (2 * sigma ** 2))
    return result
