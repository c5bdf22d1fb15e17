"""The peer-mailbox exchange protocol of csrc/peer.cuh (SyncBN statistics over NVLink peer memory) simulated on the
host: tests/peer_sim.cpp runs `world` threads as ranks over shared 64-bit atomics with the kernel's own slot / offset
arithmetic and random skew. Every rank must obtain the rank-ordered sum of every exchange (slot reuse never exposes a
stale or a future word) and nobody may wait forever -- for 2, 3, 4 and 8 ranks. (The GPU side of the same protocol is
tests/test_gpu_ddp.py on two GPUs.)"""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("peer_sim") / "peer_sim")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(HERE, "peer_sim.cpp")], check=True)
    return exe


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_rank_ordered_sums_and_no_deadlock(sim, world):
    r = subprocess.run([sim, str(world), "1500", "48", str(world)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.startswith("ok: world %d" % world)
