"""Process-wide access to the kernel library: one ``Ops`` per CUDA device.

There is deliberately no CPU implementation behind this: without the CUDA library or without a CUDA device
``get_ops`` raises. (``_TEST_OPS`` lets the CPU unit tests of the host logic inject the host-emulation library;
nothing in the package sets it.)
"""
import torch

from .kernels import Ops

_OPS = {}
_TEST_OPS = None


def get_ops(device=None) -> Ops:
    if _TEST_OPS is not None:
        return _TEST_OPS
    if device is None or torch.device(device).type != "cuda":
        if not torch.cuda.is_available():
            raise RuntimeError("mmhand_b200 runs on CUDA (sm_100a) only: no CUDA device is visible and there is no "
                               "CPU fallback")
        idx = torch.cuda.current_device()
    else:
        d = torch.device(device)
        idx = d.index if d.index is not None else torch.cuda.current_device()
    if idx not in _OPS:
        _OPS[idx] = Ops.cuda(idx)
    return _OPS[idx]


class World:
    """Data-parallel group (one process per GPU, torch.distributed).

    On CUDA the BatchNorm statistics are exchanged inside the BN finalise kernels through mailboxes in NVLink peer
    memory (``peer``: an MmhPeer handle, include/mmhand_sm100.h; csrc/peer.cu); ``MMH_SYNCBN=nccl`` keeps one NCCL
    all-reduce per exchange instead (also the path of the gloo CPU tests). Gradients are all-reduced with NCCL."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.size > 1 else 0
        self.peer = None
        self.seq = 0
        self._lib = None

    def all_reduce(self, t):
        if self.size > 1:
            self.dist.all_reduce(t)

    # -- bucketed gradient all-reduce, overlapped with the backward pass ------------------------------
    # The generator's weight gradients complete block by block on the weight-gradient side stream while the backward
    # pass goes on; each finished bucket (~33 MB of packed fp32 gradients) is summed over the ranks on a communication
    # stream of its own, so that the 285 MB all-reduce is spread under ~25 ms of backward compute instead of following
    # it (the reference's apex DDP is built with delay_allreduce=True: no overlap at all, models/MMHandModel.py:110-116).
    def comm_stream(self, device):
        if getattr(self, "_comm", None) is None:
            self._comm = torch.cuda.Stream(device)
            self._comm_ev = [torch.cuda.Event() for _ in range(2)]
        return self._comm

    def all_reduce_bucket(self, ops, view):
        """Sum ``view`` over the ranks once everything enqueued so far on the weight-gradient stream has finished."""
        if self.size <= 1:
            return
        if ops.device.type != "cuda":
            self.dist.all_reduce(view)                       # gloo (CPU tests): same collective sequence, no streams
            return
        comm = self.comm_stream(ops.device)
        producer = ops.side_stream if ops.side_stream is not None else torch.cuda.current_stream(ops.device)
        ev = self._comm_ev[0]
        ev.record(producer)
        comm.wait_event(ev)
        with torch.cuda.stream(comm):
            self.dist.all_reduce(view)
        self.buckets_issued = getattr(self, "buckets_issued", 0) + 1

    def wait_buckets(self, ops):
        """The current launch stream waits for every bucket all-reduce issued so far."""
        if self.size <= 1 or ops.device.type != "cuda" or getattr(self, "_comm", None) is None:
            return
        ev = self._comm_ev[1]
        ev.record(self._comm)
        cur = ops.in_side if ops.in_side is not None else torch.cuda.current_stream(ops.device)
        cur.wait_event(ev)

    def next_seq(self):
        self.seq = self.seq % 0xFFFFFFFF + 1          # 1, 2, ..., never 0 (the mailboxes' initial content)
        return self.seq

    def enable_peer(self, ops):
        """Create and connect the peer mailboxes (collective: every rank calls it at the same point). All ranks end
        up on the same path: if any of them cannot map its peers, everybody falls back to NCCL exchanges."""
        import ctypes
        import os
        import warnings
        if self.size <= 1 or self.peer is not None or ops.device.type != "cuda":
            return self.peer is not None
        self.tame_launches(ops)
        mode = os.environ.get("MMH_SYNCBN", "auto")
        if mode == "nccl" or not ops.lib.mmh_is_device_build():
            return False
        # (validated on 2 and 4 GPUs against the oracle on the joint batch -- tests/test_gpu_ddp.py, profiles/r02_ddp_* --
        #  and timed against per-layer NCCL all-reduces: 57.4 vs 59.0 ms per step at 4 GPUs)
        from . import lib as L
        lib, ok, handle, why = ops.lib, 1, ctypes.c_void_p(), ""
        mine = ctypes.create_string_buffer(L.PEER_HANDLE_BYTES)
        if self.size > 8 or lib.mmh_peer_create(self.rank, self.size, ctypes.byref(handle)) != 0 or \
                lib.mmh_peer_handle(handle, mine) != 0:
            ok, why = 0, lib.mmh_last_error().decode("utf-8", "replace")
        everyone = [None] * self.size
        self.dist.all_gather_object(everyone, (ok, mine.raw))
        if ok and all(o for o, _ in everyone):
            if lib.mmh_peer_connect(handle, b"".join(h for _, h in everyone)) != 0:
                ok, why = 0, lib.mmh_last_error().decode("utf-8", "replace")
        flag = torch.tensor([ok if all(o for o, _ in everyone) else 0], device=ops.device, dtype=torch.int32)
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN)
        if int(flag.item()) == 1:
            self.peer, self._lib = handle, lib
            return True
        if handle:
            lib.mmh_peer_destroy(handle)
        if self.rank == 0:
            warnings.warn("mmhand_b200: peer-memory SyncBN exchange unavailable (%s); using NCCL all-reduces" % why)
        return False

    def tame_launches(self, ops):
        """Data-parallel groups launch without programmatic dependent launch (and MMHandModel keeps the generator update
        on the launch stream): with both on, 8 GPUs stopped inside the first timed steps -- every rank blocked in a
        kernel launch with its queue full (gpurun_out/bench_n8_default.err, round 2; the same picture as round 1's
        8-GPU attempt) -- while 4 GPUs ran. CTAs parked by a dependent launch hold their SM's shared memory next to a
        BatchNorm kernel that spins on its peers' mailboxes; NCCL's kernels of the bucketed all-reduce, which need a
        few CTAs on EVERY rank to advance, then wait for an SM on one rank while another rank's exchange waits for
        that rank's statistics. Without the parked CTAs the same run takes 56.5 ms per step on 8 GPUs (2266 images/s).
        The gain given up is 0.4 ms of 57.9 at 4 GPUs. MMH_PDL=1 / MMH_G_UPDATE_STREAM=1 force them back on."""
        import os
        if self.size > 1 and ops.device.type == "cuda" and "MMH_PDL" not in os.environ and hasattr(ops.lib, "mmh_set_pdl"):
            ops.lib.mmh_set_pdl(0)

    def close(self):
        """Unmap the peers' mailboxes and free this rank's (collective in spirit: call it on every rank once no
        exchange kernel is in flight, e.g. after a barrier)."""
        if self.peer is not None:
            self._lib.mmh_peer_destroy(self.peer)
            self.peer = None

    def check(self):
        """Raise if a peer exchange timed out (a rank died): called once per optimisation step, no device sync."""
        if self.peer is not None and self._lib.mmh_peer_status(self.peer) != 0:
            raise RuntimeError("mmhand_b200: a SyncBN peer exchange timed out (another rank is gone)")
