"""Test-only: build and load the HOST EMULATION of the C ABI (libmmhand_hostemu.so).

The emulation compiles the same elementwise kernel bodies for the CPU and replaces the tcgen05 conv /
wgrad kernels by straight loops implementing their documented contract (mmhand_b200/csrc/emu_conv.cpp).
It exists so that the host logic can be tested without a GPU; the product never loads it.
"""
import glob
import os
import subprocess

from mmhand_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mmhand_b200", "csrc")
OUT = os.path.join(ROOT, "tests", "_hostemu", "libmmhand_hostemu.so")
OUT_F32 = os.path.join(ROOT, "tests", "_hostemu", "libmmhand_hostemu_f32.so")

# sources compiled in host mode: the dual-mode .cu files (as C++) and the emulation-only .cpp files
DUAL = ["api.cu", "elementwise.cu", "loss.cu", "optim.cu", "raster.cu", "peer.cu", "jointsmap.cu", "aug.cu", "input.cu", "ssim.cu"]
EMU_ONLY = ["emu_conv.cpp"]


def build(force=False, f32=False):
    """f32=True: activations stored as fp32 instead of bf16 (-DMMH_EMU_F32) -- removes rounding from the
    comparison with the fp32 oracle so that the tests check index arithmetic and calculus tightly."""
    srcs = [os.path.join(CSRC, s) for s in DUAL + EMU_ONLY if os.path.exists(os.path.join(CSRC, s))]
    deps = srcs + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(ROOT, "include", "mmhand_sm100.h")]
    out = OUT_F32 if f32 else OUT
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DMMH_HOST_EMU", "-o", out]
    if f32:
        cmd.append("-DMMH_EMU_F32")
    for s in srcs:
        cmd += ["-x", "c++", s]
    subprocess.run(cmd, check=True)
    return out


_cached = {}


def load(f32=False):
    if f32 not in _cached:
        lib = L.load(build(f32=f32))
        assert lib.mmh_is_device_build() == 0
        assert lib.act_bytes == (4 if f32 else 2)
        _cached[f32] = lib
    return _cached[f32]


def ops(f32=False):
    from mmhand_b200.kernels import Ops
    return Ops(load(f32), "cpu", lambda: 0)
