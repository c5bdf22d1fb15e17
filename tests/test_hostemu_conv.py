"""Tap tables, grid layouts and the conv / wgrad contract, checked on the host emulation against torch convs."""
import os
import subprocess
import sys

import pytest

FAST = ["gemm_64", "gemm_c16", "gemm_c32", "gemm_c48", "fwd_s1_3x3", "fwd_s1_7x7_c3", "fwd_s1_7x7_c42",
        "fwd_s1_7x7_out3", "fwd_vgg1", "fwd_s2", "fwd_up", "dgrad_s1_3x3", "dgrad_s1_7x7_c24", "dgrad_s1_7x7_out3",
        "dgrad_s2", "dgrad_up", "wgrad_s1_3x3", "wgrad_s1_7x7_c3", "wgrad_s1_7x7_out3", "wgrad_s2", "wgrad_up"]


def test_conv_cases_on_host_emulation():
    env = dict(os.environ, MMH_TEST_HOSTEMU="1")
    code = ("import sys, json; sys.path.insert(0, %r); sys.path.insert(0, %r); import conv_cases as c\n"
            "bad = [n for n in %r if not c.CASES[n]()['ok']]\nprint('BAD', bad)\nsys.exit(1 if bad else 0)"
            % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)),
               FAST))
    p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
