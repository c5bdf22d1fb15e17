"""ORACLE (test infrastructure): CPU restatement of the reference's depth-ordered hand part map
``generate_jointsmap`` (data/generic_dataset.py:30-78).

The reference draws, per bone, ``cv2.ellipse2Poly((int(mX), int(mY)), (int(length / 2), 5), int(angle), 0, 360, 1)``
and ``cv2.fillConvexPoly(temp, polygon, [avg_depth] * 3)`` on a float64 canvas, keeps the running minimum depth and
colours every pixel whose running minimum equals this bone's depth (:60-76). The reference function itself runs here
(oracle/ref_shims.py::load_reference_dataset_module) and produced tests/golden/jointsmap.npz
(oracle/make_golden_raster.py); it cannot travel to the GPU box, so there are two restatements:

* ``generate_jointsmap_cv2``  -- the reference's lines with the real OpenCV calls (cv2 4.13 in this image; the
  reference pins opencv-python 4.2.0.34) and ``np.math`` replaced by ``math``: the ground truth available here;
* ``generate_jointsmap``      -- the same with OpenCV's two routines restated in integer arithmetic
  (``ellipse2poly``: modules/imgproc/src/drawing.cpp ellipse2Poly / SinTable / cvRound; ``fill_convex_poly``:
  FillConvexPoly = outline with 8-connected LineIterator lines (clipLine) + XY_SHIFT=16 fixed-point scan
  conversion): what the CUDA kernel implements. tests/test_jointsmap.py pins it to the real cv2 calls on thousands
  of random and adversarial polygons and to the golden vectors of the reference function (with the OpenCV of this
  image, 4.13; the reference pins 4.2.0.34).
"""
import math
import sys

import numpy as np

BONES = (
    ((0, 17), 160), ((0, 1), 170), ((0, 5), 180), ((0, 9), 190), ((0, 13), 200),
    ((17, 18), 130), ((18, 19), 140), ((19, 20), 150),
    ((1, 2), 10), ((2, 3), 20), ((3, 4), 30),
    ((5, 6), 40), ((6, 7), 50), ((7, 8), 60),
    ((9, 10), 70), ((10, 11), 80), ((11, 12), 90),
    ((13, 14), 100), ((14, 15), 110), ((15, 16), 120),
)
RADIUS = 5
XY_SHIFT = 16
XY_ONE = 1 << XY_SHIFT

# OpenCV's SinTable: sin(k degrees), k = 0..450, written with seven decimals and stored as float
SIN_TABLE = np.array([float("%.7f" % math.sin(math.radians(k))) for k in range(451)], dtype=np.float32)


def bone_params(uv, depth, a, b):
    """generic_dataset.py:56-70: centre, axes, angle (all truncated with int()) and the bone's depth."""
    x0, y0 = float(uv[a][0]), float(uv[a][1])
    x1, y1 = float(uv[b][0]), float(uv[b][1])
    avg_depth = (float(depth[a]) + float(depth[b])) / 2
    mX = np.stack([x0, x1]).mean()
    mY = np.stack([y0, y1]).mean()
    length = ((x0 - x1) ** 2 + (y0 - y1) ** 2) ** 0.5
    angle = math.degrees(math.atan2(y0 - y1, x0 - x1))
    return (int(mX), int(mY)), (int(length / 2), RADIUS), int(angle), avg_depth


def generate_jointsmap_cv2(uv_coord, depth, width, height, channel=3):
    """The reference function, line by line, on the real OpenCV."""
    import cv2
    canvas = np.ones((height, width, channel)) * sys.maxsize
    _canvas = canvas.copy()
    for (a, b), color in BONES:
        temp_canvas = np.ones(canvas.shape) * sys.maxsize
        center, axes, angle, avg_depth = bone_params(uv_coord, depth, a, b)
        polygon = cv2.ellipse2Poly(center, axes, angle, 0, 360, 1)
        cv2.fillConvexPoly(temp_canvas, polygon, [avg_depth] * channel)
        _canvas = np.minimum(_canvas, temp_canvas)
        canvas[_canvas == avg_depth] = color
    canvas[canvas == sys.maxsize] = 0
    return canvas


# ------------------------------------------------------------------------------------------------ OpenCV restated
def cv_round(v):
    """cvRound(double): round half to even (lrint in the default rounding mode)."""
    return int(np.rint(v))


def ellipse2poly(center, axes, angle, arc_start=0, arc_end=360, delta=1):
    """cv::ellipse2Poly(Point, Size, int, int, int, int): double-precision points from the float sine table, then
    cvRound and removal of consecutive duplicates; a single point becomes two copies of the centre."""
    while angle < 0:
        angle += 360
    while angle > 360:
        angle -= 360
    if arc_start > arc_end:
        arc_start, arc_end = arc_end, arc_start
    while arc_start < 0:
        arc_start += 360
        arc_end += 360
    while arc_end > 360:
        arc_end -= 360
        arc_start -= 360
    if arc_end - arc_start > 360:
        arc_start, arc_end = 0, 360
    beta = SIN_TABLE[angle]                 # float
    alpha = SIN_TABLE[450 - angle]
    cx, cy = float(center[0]), float(center[1])
    w, h = float(axes[0]), float(axes[1])
    pts = []
    prev = None
    i = arc_start
    while i < arc_end + delta:
        a = min(i, arc_end)
        if a < 0:
            a += 360
        x = w * float(SIN_TABLE[450 - a])           # double = double * float
        y = h * float(SIN_TABLE[a])
        px = cx + x * float(alpha) - y * float(beta)
        py = cy + x * float(beta) + y * float(alpha)
        p = (cv_round(px), cv_round(py))
        if p != prev:
            pts.append(p)
            prev = p
        i += delta
    if len(pts) == 1:
        pts = [tuple(center), tuple(center)]
    return pts


def clip_line(width, height, p1, p2):
    """cv::clipLine(Size2l, Point2l&, Point2l&) -> (inside, p1, p2)."""
    x1, y1 = p1
    x2, y2 = p2
    right, bottom = width - 1, height - 1
    if width <= 0 or height <= 0:
        return False, p1, p2
    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8
    if (c1 & c2) == 0 and (c1 | c2) != 0:
        if c1 & 12:
            a = 0 if c1 < 8 else bottom
            x1 += int(float(a - y1) * (x2 - x1) / (y2 - y1))
            y1 = a
            c1 = (x1 < 0) + (x1 > right) * 2
        if c2 & 12:
            a = 0 if c2 < 8 else bottom
            x2 += int(float(a - y2) * (x2 - x1) / (y2 - y1))
            y2 = a
            c2 = (x2 < 0) + (x2 > right) * 2
        if (c1 & c2) == 0 and (c1 | c2) != 0:
            if c1:
                a = 0 if c1 == 1 else right
                y1 += int(float(a - x1) * (y2 - y1) / (x2 - x1))
                x1 = a
                c1 = 0
            if c2:
                a = 0 if c2 == 1 else right
                y2 += int(float(a - x2) * (y2 - y1) / (x2 - x1))
                x2 = a
                c2 = 0
    return (c1 | c2) == 0, (x1, y1), (x2, y2)


def line_pixels(width, height, pt1, pt2):
    """cv::LineIterator(img, pt1, pt2, connectivity 8, leftToRight=true): the pixels Line() writes."""
    inside = 0 <= pt1[0] < width and 0 <= pt2[0] < width and 0 <= pt1[1] < height and 0 <= pt2[1] < height
    if not inside:
        ok, pt1, pt2 = clip_line(width, height, pt1, pt2)
        if not ok:
            return []
    dx, dy = pt2[0] - pt1[0], pt2[1] - pt1[1]
    delta_x = delta_y = 1
    if dx < 0:                      # leftToRight: start from the other end
        dx, dy = -dx, -dy
        pt1 = pt2
    if dy < 0:
        dy, delta_y = -dy, -1
    vert = dy > dx
    if vert:
        dx, dy = dy, dx
        delta_x, delta_y = delta_y, delta_x
    err = dx - (dy + dy)
    plus_delta, minus_delta = dx + dx, -(dy + dy)
    minus_shift, plus_shift, minus_step, plus_step = delta_x, 0, 0, delta_y
    count = dx + 1
    if vert:
        plus_step, plus_shift = plus_shift, plus_step
        minus_step, minus_shift = minus_shift, minus_step
    x, y = pt1
    out = []
    for _ in range(count):
        out.append((x, y))
        mask = err < 0
        err += minus_delta + (plus_delta if mask else 0)
        x += minus_shift + (plus_shift if mask else 0)
        y += minus_step + (plus_step if mask else 0)
    return out


def fill_convex_poly_mask(width, height, pts):
    """cv::fillConvexPoly(img, pts, color, LINE_8, shift 0) -> boolean mask of the pixels written."""
    mask = np.zeros((height, width), dtype=bool)
    n = len(pts)
    if n == 0:
        return mask
    # outline
    p0 = pts[n - 1]
    xmin = xmax = pts[0][0]
    ymin = ymax = pts[0][1]
    imin = 0
    for i, p in enumerate(pts):
        if p[1] < ymin:
            ymin, imin = p[1], i
        ymax = max(ymax, p[1])
        xmax = max(xmax, p[0])
        xmin = min(xmin, p[0])
        for (x, y) in line_pixels(width, height, p0, p):
            mask[y, x] = True
        p0 = p
    if n < 3 or xmax < 0 or ymax < 0 or xmin >= width or ymin >= height:
        return mask
    ymax = min(ymax, height - 1)
    # scan conversion
    idx = [imin, imin]
    ye = [ymin, ymin]
    di = [1, n - 1]
    ex = [-XY_ONE, -XY_ONE]
    edx = [0, 0]
    edges = n
    y = ymin
    delta1 = delta2 = XY_ONE >> 1
    while True:
        for i in range(2):
            if y >= ye[i]:
                idx0 = idx[i]
                j = idx0 + di[i]
                if j >= n:
                    j -= n
                while True:
                    edges -= 1
                    if edges < 0:           # `for (; edges-- > 0; )` ran out
                        break
                    ty = pts[j][1]
                    if ty > y:
                        xs = pts[idx0][0] << XY_SHIFT
                        xe = pts[j][0] << XY_SHIFT
                        ye[i] = ty
                        num, den = (xe - xs) * 2 + (ty - y), 2 * (ty - y)
                        q = abs(num) // den                 # C++ integer division truncates toward zero
                        edx[i] = q if num >= 0 else -q
                        ex[i] = xs
                        idx[i] = j
                        break
                    idx0 = j
                    j += di[i]
                    if j >= n:
                        j -= n
        if edges < 0:
            break
        if y >= 0:
            left, right = (1, 0) if ex[0] > ex[1] else (0, 1)
            xx1 = (ex[left] + delta1) >> XY_SHIFT
            xx2 = (ex[right] + delta2) >> XY_SHIFT
            if xx2 >= 0 and xx1 < width:
                xx1 = max(xx1, 0)
                xx2 = min(xx2, width - 1)
                if xx2 >= xx1:
                    mask[y, xx1:xx2 + 1] = True
        ex[0] += edx[0]
        ex[1] += edx[1]
        y += 1
        if y > ymax:
            break
    return mask


def generate_jointsmap(uv_coord, depth, width, height, channel=3):
    """generate_jointsmap with OpenCV restated (the CUDA kernel's arithmetic): float64 [H, W, channel]."""
    big = float(sys.maxsize)
    run_min = np.full((height, width), big)
    canvas = np.full((height, width), big)
    for (a, b), color in BONES:
        center, axes, angle, avg_depth = bone_params(uv_coord, depth, a, b)
        mask = fill_convex_poly_mask(width, height, ellipse2poly(center, axes, angle))
        run_min = np.minimum(run_min, np.where(mask, avg_depth, big))
        canvas[run_min == avg_depth] = color
    canvas[canvas == big] = 0
    return np.repeat(canvas[:, :, None], channel, axis=2)
