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
