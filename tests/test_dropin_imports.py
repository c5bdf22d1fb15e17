"""Drop-in import surface: with this repository in front of the reference tree on sys.path,
the reference's entry points resolve models.* / losses.* / util.image_pool to this repository and everything else
(options.*, util.visualizer, util.util, data.*) to the reference. Needs the reference tree (build container only)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MMH_REFERENCE_ROOT", "/root/reference")

PROBE = r'''
import sys, types
for name, attrs in (("skimage", {}), ("skimage.draw", {"circle": None, "line_aa": None, "polygon": None}),
                    ("dominate", {}), ("dominate.tags", {}), ("visdom", {})):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
sys.modules["dominate.tags"].__dict__.update({k: None for k in ("meta", "h3", "table", "tr", "td", "p", "a", "img", "br")})
import models.MMHandModel, models.Generator, models.Discriminator, models.network_utils, models.base_model
import losses.L1_plus_perceptualLoss, util.image_pool
import util.util, util.visualizer, options.train_options
mine = [models.MMHandModel, models.Generator, models.Discriminator, models.network_utils, models.base_model,
        losses.L1_plus_perceptualLoss, util.image_pool]
theirs = [util.util, util.visualizer, options.train_options]
print("MINE", all(m.__file__.startswith(sys.argv[1]) for m in mine))
print("THEIRS", all(m.__file__.startswith(sys.argv[2]) for m in theirs))
print("VIS", hasattr(util.visualizer, "Visualizer"), hasattr(util.util, "tensor2im"))
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "util")), reason="reference tree not present")
def test_reference_entry_points_resolve_the_right_modules():
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + REF)
    r = subprocess.run([sys.executable, "-c", PROBE, ROOT, REF], capture_output=True, text=True, timeout=300, env=env,
                       cwd="/tmp")
    assert r.returncode == 0, r.stderr[-3000:]
    assert "MINE True" in r.stdout and "THEIRS True" in r.stdout and "VIS True True" in r.stdout, r.stdout


VISUALS = r'''
import sys, types, random
import numpy as np, torch
for name, attrs in (("skimage", {}), ("skimage.draw", {"circle": None, "line_aa": None, "polygon": None})):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
if not hasattr(np, "bool"):
    np.bool = bool                       # util/util.py:134 predates numpy 1.24
import hostemu
from mmhand_b200 import runtime
from mmhand_b200.options import make_opt
from mmhand_b200.rasterize import get_heatmaps
runtime._TEST_OPS = hostemu.ops(f32=True)
from models.MMHandModel import MMHandModel
torch.manual_seed(1); random.seed(1)
m = MMHandModel(make_opt(batchSize=1, fineSize=64, ngf=16, ndf=16, pool_size=0, local_rank='cpu', seed=3))
rng = np.random.RandomState(0)
g = torch.Generator().manual_seed(2)
r = lambda *s: torch.rand(*s, generator=g)
b = dict(H1=r(1, 3, 64, 64) * 2 - 1, D1=r(1, 3, 64, 64) * 2 - 1, H2=r(1, 3, 64, 64) * 2 - 1, D2=r(1, 3, 64, 64) * 2 - 1,
         P1=get_heatmaps(torch.from_numpy(rng.uniform(8, 56, (1, 21, 2))), (64, 64)),
         P2=get_heatmaps(torch.from_numpy(rng.uniform(8, 56, (1, 21, 2))), (64, 64)))
m.set_input(b)
m.optimize_parameters()
vis = m.get_current_visuals()["vis"]
pose = vis[:, 64:128]
print("SHAPE", vis.shape, vis.dtype)
print("POSE_ROWS_DIFFER", bool((pose != pose[0:1]).any()), "POSE_NONZERO", int((pose > 0).sum()) > 0)
print("FAKE_PANEL", bool((vis[:, 384:] > 0).any()))
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "util")), reason="reference tree not present")
def test_get_current_visuals_with_the_reference_drawing_code():
    """train.py:33-36: get_current_visuals() through the reference's own util.util (tensor2im, draw_pose_from_map) on a
    host-emulated model: seven panels, the pose panels are real drawings (not one row repeated)."""
    env = dict(os.environ, PYTHONPATH=os.pathsep.join((ROOT, os.path.join(ROOT, "tests"), REF)))
    r = subprocess.run([sys.executable, "-c", VISUALS], capture_output=True, text=True, timeout=600, env=env, cwd="/tmp")
    assert r.returncode == 0, r.stderr[-3000:]
    assert "SHAPE (64, 448, 3) uint8" in r.stdout, r.stdout
    assert "POSE_ROWS_DIFFER True POSE_NONZERO True" in r.stdout and "FAKE_PANEL True" in r.stdout, r.stdout
