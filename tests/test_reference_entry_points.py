"""The reference's own entry points, UNMODIFIED, on this repository's models:
``train.py`` (train.py:10-65: options, loader, MMHandModel, two iterations per epoch, visuals, error log, save) with
``--distributed`` at world size 1 (the only way its option parser survives) on a tiny RHD-shaped dataset on
disk, then ``aug.py`` (aug.py:12-71) on the checkpoint that run wrote. On a box without CUDA the kernels are the host
emulation (tests/run_reference_entry.py); under ``-m gpu`` the same two runs use the CUDA library."""
import os
import shutil
import subprocess
import sys
import tempfile

import pytest

from synth_dataset import make_rhd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MMH_REFERENCE_ROOT", "/root/reference")
RUNNER = os.path.join(ROOT, "tests", "run_reference_entry.py")

needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "train.py")), reason="reference tree not present")


def _run_both(size, ngf):
    work = tempfile.mkdtemp(prefix="mmh_entry_")          # no 'test' in the path (generic_dataset.py:114)
    try:
        data = make_rhd(os.path.join(work, "rhd"), n=8, size=size)
        env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + REF)
        args = ["--dataroot", data, "--dataset", "rhd", "--augmentation_ratio", "0.5", "--batchSize", "2",
                "--name", "exp", "--checkpoints_dir", os.path.join(work, "checkpoints"), "--niter", "1",
                "--niter_decay", "0", "--distributed", "--fineSize", str(size), "--ngf", str(ngf), "--ndf", str(ngf),
                "--nThreads", "0", "--pool_size", "4", "--display_freq", "2", "--print_freq", "2",
                "--save_latest_freq", "4", "--no_html", "--no_lsgan"]
        r = subprocess.run([sys.executable, RUNNER, "train"] + args, capture_output=True, text=True, timeout=1500,
                           env=env, cwd=work)
        assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
        ck = os.path.join(work, "checkpoints", "exp")
        files = sorted(os.listdir(ck))
        for f in ("latest_net_netG.pth", "latest_net_netD_PB.pth", "latest_net_netD_PP.pth", "1_net_netG.pth",
                  "loss_log.txt", "opt.txt"):
            assert f in files, (f, files)
        assert "End of epoch 1 / 1" in r.stdout and "saving the latest model" in r.stdout
        log = open(os.path.join(ck, "loss_log.txt")).read()
        assert "pair_L1loss" in log and "D_PP" in log and "perceptual" in log        # print_current_errors ran
        if ngf != 64:
            return                # aug.py hard-codes ngf=64 (aug.py:30-37): it is run on the full-size checkpoint only
        dst = os.path.join(work, "out")
        r = subprocess.run([sys.executable, RUNNER, "aug", "exp", data, dst, "rhd", "0.5", "0"], capture_output=True,
                           text=True, timeout=1500, env=env, cwd=work)
        assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
        import cv2
        pngs = sorted(os.listdir(os.path.join(dst, "color")))
        assert len(pngs) == 4                                   # the augmentation split: first 1 - ratio share of 8
        img = cv2.imread(os.path.join(dst, "color", pngs[0]))
        assert img.shape == (size, size, 3) and img.std() > 0
    finally:
        shutil.rmtree(work, ignore_errors=True)


@needs_ref
@pytest.mark.timeout(1800)
def test_train_py_runs_unchanged_on_the_host_emulation():
    import torch
    if torch.cuda.is_available():
        pytest.skip("covered by the gpu-marked test on this box")
    _run_both(size=32, ngf=16)


@needs_ref
@pytest.mark.gpu
@pytest.mark.timeout(1800)
def test_train_py_and_aug_py_run_unchanged_on_the_gpu():
    _run_both(size=64, ngf=64)
