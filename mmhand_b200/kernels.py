"""Python-side wrappers of the C-ABI kernels (include/mmhand_sm100.h): torch tensors in, ctypes structs out.

``Ops`` binds a loaded library and a stream getter. The product constructs it with the CUDA library
(``Ops.cuda()``); CPU tests of the host logic pass the host-emulation library explicitly.

Launch tapes. A training step is ~1.7k kernel launches whose arguments (buffer pointers, layouts) never change
once the engines are built, so building the ctypes descriptors again every step would make the step host-bound.
``Ops.record()`` captures every C call of a code region as ``(function, argument tuple)``; ``Tape.replay()`` re-issues
them with a tight loop. The few step-dependent scalars (dropout keys, Adam bias correction / lr) are refreshed by
small patch callbacks registered at record time.
"""
import ctypes as C
from contextlib import contextmanager

import torch

from . import lib as L
from .layouts import Lay

M32 = 0xFFFFFFFF


def clay(l: Lay) -> L.Lay:
    return L.Lay(l.B, l.H, l.W, l.Hg, l.Wg, l.h0, l.w0, 1 if l.phase else 0, l.ld, l.c0, l.C, 0)


def _p(t):
    return None if t is None else t.data_ptr()


def dropout_key(seed, layer_id, step):
    """Same key schedule as oracle/patn_ref.py::dropout_key."""
    return (seed * 0x9E3779B1 + layer_id * 0x85EBCA77 + step * 0xC2B2AE3D + 0x27D4EB2F) & M32


class KeyRef:
    """Dropout key of one layer as a function of the training step (resolved at launch / replay time).

    ``first_word``: hash-word index of this rank's first element in the *global* batch (data parallel: rank * local
    batch * H * W * ceil(C/8)). The kernels hash ``word * 0x9E3779B1 + key`` with the local word index, so adding
    ``first_word * 0x9E3779B1`` to the key gives exactly the mask the joint batch would get: N ranks drop the same
    elements as one process on the joint batch (and not the same pattern on every rank's own samples)."""

    def __init__(self, seed, layer_id, first_word=0):
        self.seed, self.layer_id, self.first_word = seed, layer_id, first_word

    def resolve(self, step):
        return (dropout_key(self.seed, self.layer_id, step) + self.first_word * 0x9E3779B1) & M32


class GradSource:
    """Data gradient of a consumer convolution, addressed in the consumer's input layout."""

    def __init__(self, buf, lay: Lay, pad_lo, pad_hi, reflect):
        self.buf, self.lay, self.pad_lo, self.pad_hi, self.reflect = buf, lay, pad_lo, pad_hi, reflect

    def view(self, c0, Cv):
        return GradSource(self.buf, self.lay.view(c0, Cv), self.pad_lo, self.pad_hi, self.reflect)

    def c(self):
        s = L.GradSrc()
        s.p = self.buf.data_ptr()
        s.l = clay(self.lay)
        s.pad_lo, s.pad_hi, s.reflect = self.pad_lo, self.pad_hi, 1 if self.reflect else 0
        return s


class Tape:
    """Recorded launch sequence. Entries: [fn, args(list), patch] with fn None for host callbacks."""

    def __init__(self, ops, stream):
        self.ops, self.stream, self.cmds, self.keep = ops, stream, [], []

    def replay(self, step):
        ops = self.ops
        ops.step = step
        err = 0
        for fn, args, patch in self.cmds:
            if patch is not None:
                patch(step, args)
            if fn is None:
                args()
            else:
                err |= fn(*args)
        ops.launches += self.n_launches
        if err:
            raise L.MmhError(ops.lib.mmh_last_error().decode("utf-8", "replace"))

    @property
    def n_launches(self):
        return sum(1 for c in self.cmds if c[0] is not None and not getattr(c[0], "_mmh_no_kernel", False))


class Ops:
    def __init__(self, lib, device, stream_fn):
        self.lib = lib
        self.device = torch.device(device)
        self._stream = stream_fn
        self.launches = 0
        self.act_dtype = torch.bfloat16 if lib.act_bytes == 2 else torch.float32
        self.tape = None
        self.step = 0            # training step used to resolve KeyRefs in immediate mode
        self.main_stream = None  # optional high-priority stream of the critical path (see Ops.cuda)
        self.in_side = None      # the torch stream launches currently go to, when it is a side stream
        self.side_stream = None  # torch.cuda.Stream of the weight-gradient kernels (None: everything on one stream)
        self.chains = [None]     # side-stream index per layer chain (one entry: no chain concurrency)
        self._ev = None
        self.conv_hook = None    # optional callable(kind, plan, launch) wrapping conv launches (bench instrumentation)
        self.ew_hook = None      # optional callable(kernel class, algorithmic bytes, launch) for the bandwidth-bound kernels

    @staticmethod
    def cuda(device=None):
        lib = L.load()
        if not torch.cuda.is_available():
            raise L.MmhError("mmhand_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        ops = Ops(lib, dev, lambda: torch.cuda.current_stream(dev).cuda_stream)
        import os
        if os.environ.get("MMH_WGRAD_STREAM", "1") != "0":
            chains = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)] \
                if os.environ.get("MMH_PAT_STREAMS", "1") != "0" else []
            ops.enable_side_stream(torch.cuda.Stream(dev), torch.cuda.Stream(dev), chains)
            if os.environ.get("MMH_MAIN_PRIORITY", "0") != "0":
                # critical-path launches on a high-priority stream: its CTAs are placed before pending side-stream ones
                ops.main_stream = torch.cuda.Stream(dev, priority=-1)
        return ops

    # ------------------------------------------------------------------ side stream (weight gradients)
    def enable_side_stream(self, stream, update_stream=None, chain_streams=()):
        """stream: weight-gradient kernels (index 0); update_stream: optimiser update of one network under the
        compute of another (index 1, defaults to the same stream); chain_streams: indices 2, 3 -- the second and third
        stream of a PAT block / stem run their layer chains there, next to the first one's on the main stream
        (``chains``: the side-stream index of chain s, None = the launch stream)."""
        self.side_stream = stream
        self._sides = [stream, update_stream if update_stream is not None else stream] + list(chain_streams)
        self.chains = [None] + [2 + i for i in range(len(chain_streams))]
        self._ev = []
        for _ in range(2 * len(self._sides)):
            e = C.c_void_p()
            if self.lib.mmh_event_create(C.byref(e)) != 0:
                raise L.MmhError(self.lib.mmh_last_error().decode("utf-8", "replace"))
            self._ev.append(e)

    def fork(self, which=0):
        """Everything enqueued on the main stream so far happens-before what is launched inside ``side(which)`` next.
        (An event may be re-recorded at once: a wait refers to the record that preceded it.)"""
        side = C.c_void_p(self._sides[which].cuda_stream)
        ev = self._ev[2 * which]
        self._run(self.lib.mmh_event_record, (ev, self.st()))
        self._run(self.lib.mmh_stream_wait_event, (side, ev))
        self.launches -= 2

    def join(self, which=0):
        """The main stream waits for everything launched on side stream ``which`` so far."""
        side = C.c_void_p(self._sides[which].cuda_stream)
        ev = self._ev[2 * which + 1]
        self._run(self.lib.mmh_event_record, (ev, side))
        self._run(self.lib.mmh_stream_wait_event, (self.st(), ev))
        self.launches -= 2

    @contextmanager
    def side(self, which=0):
        """Launches inside go to side stream ``which`` (None: no change). Contexts nest: a weight gradient launched
        from inside a chain forks off the chain's stream."""
        if which is None:
            yield
            return
        old, old_in = self._stream, self.in_side
        h = self._sides[which].cuda_stream
        self._stream = lambda: h
        self.in_side = self._sides[which]
        try:
            yield
        finally:
            self._stream, self.in_side = old, old_in

    def st(self):
        return C.c_void_p(self._stream())

    def _run(self, fn, args, patch=None, keep=None, meta=None):
        """meta = (kernel class, algorithmic bytes): lets bench.py bracket bandwidth-bound launches with CUDA events
        (``ew_hook``; eager launches only, like ``conv_hook``)."""
        if meta is not None and self.ew_hook is not None and self.tape is None:
            hook, self.ew_hook = self.ew_hook, None
            try:
                hook(meta[0], meta[1], lambda: self._run(fn, args, patch, keep))
            finally:
                self.ew_hook = hook
            return
        rc = fn(*args)
        self.launches += 1
        if rc != 0:
            raise L.MmhError(self.lib.mmh_last_error().decode("utf-8", "replace"))
        if self.tape is not None:
            self.tape.cmds.append([fn, list(args), patch])
            if keep is not None:
                self.tape.keep.append(keep)

    def host(self, fn):
        """Host-side bookkeeping that must be repeated on every replay (e.g. BN batch counters)."""
        fn()
        if self.tape is not None:
            self.tape.cmds.append([None, fn, None])

    @contextmanager
    def record(self):
        assert self.tape is None, "nested tape recording"
        tape = Tape(self, self._stream())
        self.tape = tape
        try:
            yield tape
        finally:
            self.tape = None

    # Engine buffers are carved out of zero-filled chunks: an engine owns ~800 activation / gradient / statistics
    # buffers, and one allocator call + one fill kernel each made engine construction ~1700 driver calls (and buried
    # the library's own kernels under fill launches in any launch list of a first step). A chunk is released by the
    # caching allocator when the last buffer carved from it dies (views keep their storage alive).
    ARENA_CHUNK = 256 << 20
    ARENA_ALIGN = 1024               # TMA base addresses need 128 B; 1 KB keeps swizzle atoms aligned too
    # debugging aids: MMH_ARENA=0 gives every buffer its own torch allocation; MMH_ARENA_GUARD=<bytes> leaves a zero
    # gap behind every carved buffer, and Ops.check_guards() reports the buffers whose gap was written to
    import os as _os
    ARENA_ON = _os.environ.get("MMH_ARENA", "1") != "0"
    ARENA_GUARD = int(_os.environ.get("MMH_ARENA_GUARD", "0"))

    def zeros(self, *shape, dtype=None):
        dtype = dtype or self.act_dtype
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        n = 1
        for d in shape:
            n *= int(d)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        if nbytes == 0 or nbytes > self.ARENA_CHUNK // 2 or not self.ARENA_ON:
            return torch.zeros(*shape, dtype=dtype, device=self.device)
        need = (nbytes + self.ARENA_GUARD + self.ARENA_ALIGN - 1) // self.ARENA_ALIGN * self.ARENA_ALIGN
        chunk = getattr(self, "_arena", None)
        if chunk is None or self._arena_off + need > chunk.numel():
            chunk = self._arena = torch.zeros(self.ARENA_CHUNK, dtype=torch.uint8, device=self.device)
            self._arena_off = (-chunk.data_ptr()) % self.ARENA_ALIGN
        off = self._arena_off
        self._arena_off = off + need
        if self.ARENA_GUARD:
            import traceback
            who = [f for f in traceback.extract_stack(limit=6) if "kernels.py" not in f.filename]
            self.__dict__.setdefault("_guards", []).append(
                (chunk, off + nbytes, off + need, tuple(shape), str(dtype),
                 " <- ".join("%s:%d" % (f.filename.split("/")[-1], f.lineno) for f in reversed(who[-3:]))))
        return chunk[off:off + nbytes].view(dtype).view(*shape)

    def check_guards(self):
        """MMH_ARENA_GUARD debugging: the buffers whose trailing gap is no longer zero (out-of-bounds writes)."""
        bad = []
        for chunk, lo, hi, shape, dtype, who in getattr(self, "_guards", []):
            g = chunk[lo:hi]
            if bool(g.any()):
                nz = g.nonzero().flatten()
                bad.append("buffer %s %s (%s): %d dirty guard bytes, first at +%d, last at +%d" % (
                    shape, dtype, who, nz.numel(), int(nz[0]), int(nz[-1])))
        return bad

    def empty(self, *shape, dtype=None):
        dtype = dtype or self.act_dtype
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def memset0(self, t):
        self._run(self.lib.mmh_memset, (t.data_ptr(), 0, t.numel() * t.element_size(), self.st()), keep=t)

    def _key(self, key, struct):
        """Resolve a dropout key now and return the patch that refreshes it on replay."""
        if isinstance(key, KeyRef):
            struct.drop_key = key.resolve(self.step)
            return lambda step, args, s=struct, k=key: setattr(s, "drop_key", k.resolve(step))
        struct.drop_key = int(key) & M32
        return None

    # ------------------------------------------------------------------ convolutions
    def run_conv(self, plan, tag=None):
        if self.conv_hook is not None and self.tape is None:
            self.conv_hook("conv", tag, plan, lambda: self._run(self.lib.mmh_conv_run, (plan.handle, self.st())))
        else:
            self._run(self.lib.mmh_conv_run, (plan.handle, self.st()), keep=plan)

    def run_conv_key(self, plan, key, tag=None):
        """Data-gradient launch with fused BN-backward masks: ``key`` is the producer layer's dropout key (a KeyRef is
        resolved now and again on every tape replay)."""
        if isinstance(key, KeyRef):
            k0 = key.resolve(self.step)
            patch = lambda step, args, k=key: args.__setitem__(1, k.resolve(step))
        else:
            k0, patch = int(key) & M32, None
        if self.conv_hook is not None and self.tape is None:
            self.conv_hook("conv", tag, plan, lambda: self._run(self.lib.mmh_conv_run_key, (plan.handle, k0, self.st())))
        else:
            self._run(self.lib.mmh_conv_run_key, (plan.handle, k0, self.st()), patch, keep=plan)

    def run_wgrad(self, plan, tag=None):
        if self.conv_hook is not None and self.tape is None:
            self.conv_hook("wgrad", tag, plan, lambda: self._run(self.lib.mmh_wgrad_run, (plan.handle, self.st())))
        else:
            self._run(self.lib.mmh_wgrad_run, (plan.handle, self.st()), keep=plan)

    # ------------------------------------------------------------------ forward elementwise
    def assemble(self, src0, src1, dst, dl: Lay, pad_lo, pad_hi, reflect, scale=None, shift=None):
        c0 = src0.shape[1]
        c1 = src1.shape[1] if src1 is not None else 0
        cl = clay(dl)
        self._run(self.lib.mmh_assemble_nchw, (_p(src0), c0, _p(src1), c1, _p(scale), _p(shift), _p(dst), C.byref(cl),
                                               pad_lo, pad_hi, 1 if reflect else 0, self.st()), keep=(cl, src0, src1))

    def bn_stats(self, x, rows, ld, Cc, sums):
        self._run(self.lib.mmh_bn_stats, (_p(x), rows, ld, Cc, _p(sums), self.st()))

    def bn_finalize(self, sums, count, gamma, beta, rm, rv, momentum, eps, train, Cc, coef, save):
        self._run(self.lib.mmh_bn_finalize, (_p(sums), float(count), _p(gamma), _p(beta), _p(rm), _p(rv), momentum, eps,
                                             1 if train else 0, Cc, _p(coef), _p(save), self.st()))

    def _seq_args(self, world, args):
        """Arguments of a launch fused with a peer exchange: slot 1 holds the exchange's sequence number, drawn from
        the group's counter now and again on every tape replay (all ranks issue the same exchanges in order)."""
        args = [world.peer, world.next_seq()] + list(args)
        return args, (lambda step, a, w=world: a.__setitem__(1, w.next_seq()))

    def _peer_args(self, world, args, patch=None):
        """(peer, seq) prefix of the one-launch statistics kernels: no group -> (NULL, 0) and only the caller's
        patch; with a group, the sequence number is refreshed on replay as well."""
        if world is None:
            return [None, 0] + list(args), patch
        args, seq_patch = self._seq_args(world, args)
        if patch is None:
            return args, seq_patch
        return args, (lambda step, a, p0=patch, p1=seq_patch: (p0(step, a), p1(step, a)))

    def bn_stats_finalize(self, world, x, rows, ld, Cc, sums, counter, count_global, gamma, beta, rm, rv, momentum,
                          eps, coef, save):
        """Train-mode BatchNorm statistics in ONE launch: per-channel sums, (peer exchange,) mean / rstd / affine
        coefficients / running statistics in the last block, accumulators back to zero."""
        args, patch = self._peer_args(world, (_p(x), rows, ld, Cc, _p(sums), _p(counter), float(count_global),
                                              _p(gamma), _p(beta), _p(rm), _p(rv), momentum, eps, _p(coef), _p(save),
                                              self.st()))
        self._run(self.lib.mmh_bn_stats_finalize, args, patch)

    def bn_finalize_reset(self, world, sums, count_global, gamma, beta, rm, rv, momentum, eps, Cc, coef, save):
        """Finalisation of statistics accumulated by a convolution epilogue; ``sums`` returns to zero."""
        args, patch = self._peer_args(world, (_p(sums), float(count_global), _p(gamma), _p(beta), _p(rm), _p(rv),
                                              momentum, eps, Cc, _p(coef), _p(save), self.st()))
        self._run(self.lib.mmh_bn_finalize_reset, args, patch)

    def bn_bwd_reduce_finalize(self, world, dz, dz_f32, relu, dropout, key, x, xl, coef, save, sums, k, counter,
                               count_global, dgamma, dbeta):
        p, patch = self._bn_bwd(dz, dz_f32, relu, dropout, key, x, xl, coef, save, sums=sums, k=k)
        args, patch = self._peer_args(world, (C.byref(p), _p(counter), float(count_global), _p(dgamma), _p(dbeta),
                                              self.st()), patch)
        self._run(self.lib.mmh_bn_bwd_reduce_finalize, args, patch, keep=p,
                  meta=("bn_bwd", 2 * self.lib.act_bytes * xl.B * xl.H * xl.W * xl.C))

    def gate_bwd_reduce_finalize(self, world, dout, c1, x2o, x3o, sl, coef, save, sums, k, counter, count_global,
                                 dgamma, dbeta):
        p = self._gate_bwd(dout, c1, x2o, x3o, sl, coef, save, sums=sums, k=k)
        args, patch = self._peer_args(world, (C.byref(p), _p(counter), float(count_global), _p(dgamma), _p(dbeta),
                                              self.st()))
        self._run(self.lib.mmh_gate_bwd_reduce_finalize, args, patch, keep=p)

    def bn_bwd_finalize_reset(self, world, sums, count_global, k, dgamma, dbeta, Cc):
        """Finalisation of BN-backward sums accumulated by a data-gradient epilogue; ``sums`` returns to zero."""
        args, patch = self._peer_args(world, (_p(sums), float(count_global), _p(k), _p(dgamma), _p(dbeta), Cc, self.st()))
        self._run(self.lib.mmh_bn_bwd_finalize_reset, args, patch)

    def bn_finalize_sync(self, world, sums, count_global, gamma, beta, rm, rv, momentum, eps, Cc, coef, save):
        """Train-mode BN finalisation on the statistics of all ranks (exchange over NVLink peer memory inside the
        kernel; ``sums`` becomes the global sums)."""
        args, patch = self._seq_args(world, (_p(sums), float(count_global), _p(gamma), _p(beta), _p(rm), _p(rv),
                                             momentum, eps, Cc, _p(coef), _p(save), self.st()))
        self._run(self.lib.mmh_bn_finalize_sync, args, patch)

    def bn_bwd_finalize_sync(self, world, sums_local, sums_global, count_global, k, dgamma, dbeta, Cc):
        args, patch = self._seq_args(world, (_p(sums_local), _p(sums_global), float(count_global), _p(k), _p(dgamma),
                                             _p(dbeta), Cc, self.st()))
        self._run(self.lib.mmh_bn_bwd_finalize_sync, args, patch)

    def peer_sum(self, world, data):
        args, patch = self._seq_args(world, (_p(data), data.numel(), self.st()))
        self._run(self.lib.mmh_peer_sum, args, patch, keep=data)

    def norm_act(self, src, sl: Lay, coef, relu, dropout, key, dst, dl: Lay, pad_lo, pad_hi, reflect, resid=None,
                 dst_f32=None):
        p = L.NormAct()
        p.src, p.sl, p.coef = _p(src), clay(sl), _p(coef)
        p.relu, p.dropout = int(relu), int(dropout)
        patch = self._key(key, p)
        p.resid, p.dst = _p(resid), _p(dst)
        p.dl = clay(dl) if dl is not None else clay(sl)
        p.pad_lo, p.pad_hi, p.reflect = pad_lo, pad_hi, 1 if reflect else 0
        p.dst_f32 = _p(dst_f32)
        ab, n_el = self.lib.act_bytes, sl.B * sl.H * sl.W * sl.C
        out_el = (dl.rows * dl.C if (dl is not None and dst is not None) else 0)
        by = ab * n_el + ab * out_el + (4 * n_el if resid is not None else 0) + (4 * n_el if dst_f32 is not None else 0)
        self._run(self.lib.mmh_norm_act, (C.byref(p), self.st()), patch, keep=p, meta=("norm_act", by))

    def gate_fwd(self, c1, x2o, x3o, sl, coef, trunk_in, trunk_out, d1, d1l, d2, d2l, d3, d3l, pad_lo, pad_hi, reflect):
        p = L.GateFwd()
        p.c1, p.x2o, p.x3o, p.sl, p.coef = _p(c1), _p(x2o), _p(x3o), clay(sl), _p(coef)
        p.trunk_in, p.trunk_out = _p(trunk_in), _p(trunk_out)
        p.d1, p.d1l = _p(d1), clay(d1l)
        p.d2, p.d2l = _p(d2), clay(d2l if d2l is not None else d1l)
        p.d3, p.d3l = _p(d3), clay(d3l if d3l is not None else d1l)
        p.pad_lo, p.pad_hi, p.reflect = pad_lo, pad_hi, 1 if reflect else 0
        self._run(self.lib.mmh_gate_fwd, (C.byref(p), self.st()), keep=p)

    # ------------------------------------------------------------------ backward elementwise
    def grad_gather(self, srcs, B, H, W, Cc, dst, dl: Lay, dst_f32, trunk=None, mask=None, ml=None):
        p = L.GradGather()
        p.nsrc, p.dst_f32, p.B, p.H, p.W, p.C = len(srcs), 1 if dst_f32 else 0, B, H, W, Cc
        for i, s in enumerate(srcs):
            p.src[i] = s.c()
        p.trunk, p.mask = _p(trunk), _p(mask)
        p.ml = clay(ml if ml is not None else dl)
        p.dst, p.dl = _p(dst), clay(dl)
        self._run(self.lib.mmh_grad_gather, (C.byref(p), self.st()), keep=p)

    def _bn_bwd(self, dz, dz_f32, relu, dropout, key, x, xl, coef, save, sums=None, k=None, dy=None, yl=None):
        """dz: plain buffer, or a tuple (sources, trunk) = gather the upstream gradient inside the kernel."""
        p = L.BnBwd()
        if isinstance(dz, tuple):
            srcs, trunk = dz
            assert len(srcs) <= 2 and (srcs or trunk is not None)
            if not srcs:            # the fp32 trunk alone is just a plain fp32 dz
                dz, dz_f32 = trunk, True
            else:
                p.nsrc, p.trunk, dz = len(srcs), _p(trunk), None
                for i, s in enumerate(srcs):
                    p.src[i] = s.c()
        p.dz, p.dz_f32, p.relu, p.dropout = _p(dz), 1 if dz_f32 else 0, int(relu), int(dropout)
        patch = self._key(key, p)
        p.x, p.xl, p.coef, p.save = _p(x), clay(xl), _p(coef), _p(save)
        p.sums, p.k, p.dy = _p(sums), _p(k), _p(dy)
        p.yl = clay(yl if yl is not None else xl)
        return p, patch

    def bn_bwd_reduce(self, dz, dz_f32, relu, dropout, key, x, xl, coef, save, sums):
        p, patch = self._bn_bwd(dz, dz_f32, relu, dropout, key, x, xl, coef, save, sums=sums)
        self._run(self.lib.mmh_bn_bwd_reduce, (C.byref(p), self.st()), patch, keep=p)

    def bn_bwd_apply(self, dz, dz_f32, relu, dropout, key, x, xl, coef, save, k, dy, yl):
        p, patch = self._bn_bwd(dz, dz_f32, relu, dropout, key, x, xl, coef, save, k=k, dy=dy, yl=yl)
        self._run(self.lib.mmh_bn_bwd_apply, (C.byref(p), self.st()), patch, keep=p,
                  meta=("bn_bwd", 3 * self.lib.act_bytes * xl.B * xl.H * xl.W * xl.C))

    def bn_bwd_finalize(self, sums_global, sums_local, count, k, dgamma, dbeta, Cc):
        self._run(self.lib.mmh_bn_bwd_finalize, (_p(sums_global), _p(sums_local), float(count), _p(k), _p(dgamma),
                                                 _p(dbeta), Cc, self.st()))

    def _gate_bwd(self, dout, c1, x2o, x3o, sl, coef, save, sums=None, k=None, ex2=None, ex3=None, dy1=None,
                  dy2=None, dy3=None, yl=None):
        p = L.GateBwd()
        p.dout, p.c1, p.x2o, p.x3o, p.sl = _p(dout), _p(c1), _p(x2o), _p(x3o), clay(sl)
        p.coef, p.save, p.sums, p.k = _p(coef), _p(save), _p(sums), _p(k)
        if ex2 is not None:
            p.ex2 = ex2.c()
        if ex3 is not None:
            p.ex3 = ex3.c()
        p.dy1, p.dy2, p.dy3 = _p(dy1), _p(dy2), _p(dy3)
        p.yl = clay(yl if yl is not None else sl)
        return p

    def gate_bwd_reduce(self, dout, c1, x2o, x3o, sl, coef, save, sums):
        p = self._gate_bwd(dout, c1, x2o, x3o, sl, coef, save, sums=sums)
        self._run(self.lib.mmh_gate_bwd_reduce, (C.byref(p), self.st()), keep=p)

    def gate_bwd_apply(self, dout, c1, x2o, x3o, sl, coef, save, k, ex2, ex3, dy1, dy2, dy3, yl):
        p = self._gate_bwd(dout, c1, x2o, x3o, sl, coef, save, k=k, ex2=ex2, ex3=ex3, dy1=dy1, dy2=dy2, dy3=dy3, yl=yl)
        self._run(self.lib.mmh_gate_bwd_apply, (C.byref(p), self.st()), keep=p)

    # ------------------------------------------------------------------ losses
    def bce_logits(self, x, target, loss_scale, grad_scale, loss_acc, grad=None):
        self._run(self.lib.mmh_bce_logits, (_p(x), x.numel(), float(target), loss_scale, grad_scale, _p(loss_acc),
                                            _p(grad), self.st()), keep=(x, loss_acc, grad))

    def l1(self, a, b, loss_scale, grad_scale, loss_acc, grad_acc=None):
        self._run(self.lib.mmh_l1_f32, (_p(a), _p(b), a.numel(), loss_scale, grad_scale, _p(loss_acc), _p(grad_acc),
                                        self.st()), keep=(a, b, loss_acc, grad_acc))

    def perc_loss(self, ff, ft, mse, loss_scale, grad_scale, loss_acc, dy=None, linear=False):
        """linear: the features end on a convolution (no ReLU mask in the gradient)."""
        self._run(self.lib.mmh_perc_loss, (_p(ff), _p(ft), ff.numel(), (1 if mse else 0) | (2 if linear else 0), loss_scale, grad_scale,
                                           _p(loss_acc), _p(dy), self.st()), keep=(loss_acc,))

    def tanh_bwd(self, dfake, fake, dy, yl: Lay, Cc):
        cl = clay(yl)
        self._run(self.lib.mmh_tanh_bwd, (_p(dfake), _p(fake), _p(dy), C.byref(cl), Cc, self.st()), keep=(cl, dfake))

    def input_grad_nchw(self, src: GradSource, scale, dst, B, Cc, H, W, accumulate):
        s = src.c()
        self._run(self.lib.mmh_input_grad_nchw, (C.byref(s), _p(scale), _p(dst), B, Cc, H, W, 1 if accumulate else 0,
                                                 self.st()), keep=(s, dst))

    def grid_to_nchw(self, src, sl: Lay, dst, Cc):
        cl = clay(sl)
        self._run(self.lib.mmh_grid_to_nchw, (_p(src), C.byref(cl), _p(dst), Cc, self.st()), keep=cl)

    # ------------------------------------------------------------------ parameters
    def pack_weight(self, src, s_n, s_c, s_t, N, Cc, T, dst, Np, Cp):
        self._run(self.lib.mmh_pack_weight, (_p(src), s_n, s_c, s_t, N, Cc, T, _p(dst), Np, Cp, self.st()))

    def pack_weight_folded(self, src, s_n, s_c, s_t, N, Cc, kh, kw, dst, Np, Cin_p, Kw, reverse=False):
        self._run(self.lib.mmh_pack_weight_folded, (_p(src), s_n, s_c, s_t, N, Cc, kh, kw, _p(dst), Np, Cin_p, Kw,
                                                    1 if reverse else 0, self.st()))

    def unpack_wgrad(self, src, dst, s_n, s_c, s_t, N, Cc, T, accumulate):
        self._run(self.lib.mmh_unpack_wgrad, (_p(src), _p(dst), s_n, s_c, s_t, N, Cc, T, 1 if accumulate else 0,
                                              self.st()))

    def make_param_jobs(self, jobs):
        """jobs: list of dicts (kind, src, dst, s_n, s_c, N, C, T, Np, Cp) -> (device table, n, tiles, keep-alive)."""
        arr = (L.ParamJob * len(jobs))()
        tile = 0
        for a, j in zip(arr, jobs):
            a.src, a.dst = j["src"].data_ptr(), j["dst"].data_ptr()
            a.s_n, a.s_c, a.kind = j["s_n"], j["s_c"], j["kind"]
            a.N, a.C, a.T, a.Np, a.Cp = j["N"], j["C"], j["T"], j.get("Np", j["N"]), j.get("Cp", j["C"])
            a.tile_begin = tile
            pairs = a.Np * a.Cp if a.kind == 0 else a.N * a.C
            tile += (pairs + 255) // 256
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone()
        table = host.to(self.device)
        return table, len(jobs), tile, [j["src"] for j in jobs] + [j["dst"] for j in jobs]

    def run_param_jobs(self, packed):
        table, n, tiles, keep = packed
        if n:
            self._run(self.lib.mmh_param_jobs, (table.data_ptr(), n, tiles, self.st()), keep=(table, keep))

    def adam(self, p, g, m, v, lr, b1, b2, eps, step, grad_scale=1.0, dyn=None):
        """dyn: optional callable() -> (lr, step) evaluated again on every replay."""
        patch = None
        if dyn is not None:
            def patch(_step, args, dyn=dyn):
                args[5], args[9] = dyn()
        self._run(self.lib.mmh_adam, (_p(p), _p(g), _p(m), _p(v), p.numel(), lr, b1, b2, eps, step, grad_scale,
                                      self.st()), patch)

    def image_pack_bgr8(self, src, dst):
        n, _, H, W = src.shape
        self._run(self.lib.mmh_image_pack_bgr8, (_p(src), n, H, W, _p(dst), self.st()), keep=(src, dst))

    def ssim(self, a, b, window, sigma, mean_acc, per_image):
        B, Cc, H, W = a.shape
        self._run(self.lib.mmh_ssim, (_p(a), _p(b), B, Cc, H, W, window, float(sigma), _p(mean_acc), _p(per_image),
                                      self.st()), keep=(a, b, mean_acc, per_image))

    def image_unpack_u8(self, src, dst, swap_rb=False):
        n, H, W, _ = src.shape
        self._run(self.lib.mmh_image_unpack_u8, (_p(src), n, H, W, 1 if swap_rb else 0, _p(dst), self.st()),
                  keep=(src, dst))

    def depth_unpack_u8(self, src, dst, hi=1, lo=2, div=700.0):
        n, H, W, _ = src.shape
        self._run(self.lib.mmh_depth_unpack_u8, (_p(src), n, H, W, hi, lo, float(div), _p(dst), self.st()),
                  keep=(src, dst))

    def jointsmap(self, uv, depth, H, W, out_f64, out_u8):
        n = uv.numel() // 42
        self._run(self.lib.mmh_jointsmap_rasterize, (_p(uv), _p(depth), n, H, W, _p(out_f64), _p(out_u8), self.st()),
                  keep=(uv, depth, out_f64, out_u8))

    def pose_maps(self, yx, H, W, sigma, missing, out):
        n, J = yx.shape[0], yx.shape[1]
        self._run(self.lib.mmh_pose_map_rasterize, (_p(yx), n, J, H, W, float(sigma), float(missing), _p(out), self.st()),
                  keep=(yx, out))

    def heatmaps(self, uv, H, W, sigma, thresh, out):
        n = uv.numel() // 2
        self._run(self.lib.mmh_heatmap_rasterize, (_p(uv), n, H, W, float(sigma), float(thresh), _p(out), self.st()))
