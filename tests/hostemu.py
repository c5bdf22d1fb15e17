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

# sources compiled in host mode: the dual-mode .cu files (as C++) and the emulation-only .cpp files
DUAL = ["api.cu", "elementwise.cu", "loss.cu", "optim.cu", "raster.cu"]
EMU_ONLY = ["emu_conv.cpp"]


def build(force=False):
    srcs = [os.path.join(CSRC, s) for s in DUAL + EMU_ONLY if os.path.exists(os.path.join(CSRC, s))]
    deps = srcs + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(ROOT, "include", "mmhand_sm100.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-DMMH_HOST_EMU", "-o", OUT]
    for s in srcs:
        cmd += ["-x", "c++", s]
    subprocess.run(cmd, check=True)
    return OUT


_cached = None


def load():
    global _cached
    if _cached is None:
        _cached = L.load(build())
        assert _cached.mmh_is_device_build() == 0
    return _cached
