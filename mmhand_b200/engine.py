"""Execution engines of the three networks on the C-ABI kernels.

An engine owns every activation / gradient buffer of one network for one (batch, frame size), the packed
bf16 tensor-core operands of its weights and the conv plans bound to those buffers; ``forward`` and
``backward`` are plain sequences of kernel launches on the current stream (no autograd, no allocation,
no host synchronisation). Layer structure and arithmetic follow the reference modules:

  GeneratorEngine      models/Generator.py:133-283 (PATNModel) incl. PATBlock :8-130
  DiscriminatorEngine  models/Discriminator.py:58-154
  VggEngine            losses/L1_plus_perceptualLoss.py:22-27,54-61 (VGG19.features[0:4])
"""
import os

import torch

from . import convops
from .kernels import GradSource, KeyRef, Ops, dropout_key  # noqa: F401
from .layouts import Lay, chan_pad, geom_s1, geom_s2, geom_up

BN_EPS = 1e-5
BN_MOM = 0.1

# MMH_FUSE_GATHER=0 materialises dz with mmh_grad_gather before the BN backward (A/B measurements only)
FUSE_GATHER = os.environ.get("MMH_FUSE_GATHER", "1") != "0"
# BN statistics in the conv epilogue: 0 off, 1 every BatchNorm'd conv, 2 (default) only where the main loop hides the
# longer epilogue (contraction length taps * channels >= 2304: the 256/512-channel 3x3 layers; measured on the 7x7
# stems and the stride-2 layers the epilogue is the bottleneck and the fused statistics cost more than their kernel)
CONV_STATS = int(os.environ.get("MMH_CONV_STATS", "2"))
# BatchNorm-backward sums (and the ReLU / dropout masks) inside the data-gradient epilogue of the consumer convolution
# (conv_p -> BN -> ReLU -> dropout -> reflect pad -> conv_c: conv_c's dgrad stores the masked gradient and accumulates
# sum dze, sum dze * xhat of conv_p's BatchNorm): the reduction pass over dz and x disappears. Measured on B200
# (profiles/r02_bn_bwd_epilogue.txt): the fused launch takes 187 us against 105 us (plain data gradient) + 62 us
# (reduction kernel) at 512 channels -- with one epilogue warp per scheduler the ~520 instructions per 16-column chunk
# are latency-bound and the epilogue (25 us per tile) outlasts the main loop (14 us) -- so it is OFF by default
# (MMH_FUSE_BN_BWD=1 enables it; parity-tested either way).
FUSE_BN_BWD = os.environ.get("MMH_FUSE_BN_BWD", "0") != "0"
# data parallel: all-reduce the generator's weight gradients bucket by bucket during its backward pass (0: one all-reduce
# of the whole gradient after the pass)
GRAD_BUCKETS = os.environ.get("MMH_GRAD_BUCKETS", "1") != "0"
FUSE_BN_BWD_MIN_K = int(os.environ.get("MMH_FUSE_BN_BWD_MIN_K", "2304"))    # taps * channels of the contraction


def plain_lay(B, H, W, Cc):
    return Lay(B, H, W, H, W, 0, 0, False, Cc, 0, Cc)


class ParamStore:
    """Flat fp32 storage (values, grads, Adam moments, step count) behind the nn.Parameters of one module.

    Owned by the *module* (``param_store``), not by an engine: engines are per (batch, frame size) and are rebuilt when
    the batch shape changes (the ragged last batch of an epoch -- the reference's DataLoader has no drop_last), while
    the optimiser state must survive, as torch.optim.Adam's does in the reference (models/MMHandModel.py:90-98)."""

    def __init__(self, ops: Ops, module):
        self.ops, self.module = ops, module
        # flat order: BatchNorm scales / shifts and biases first (one small contiguous region: in data-parallel runs
        # with bucketed weight-gradient all-reduces it is the only part of ``grad`` that still has to be summed over
        # the ranks at the end of a backward pass), then the convolution weights
        ps = [p for p in module.parameters()]
        self.params = [p for p in ps if p.dim() != 4] + [p for p in ps if p.dim() == 4]
        self.n_small = sum(p.numel() for p in ps if p.dim() != 4)
        self.flat = None
        self.step = 0
        self.m = self.v = None
        self.ensure()

    def ensure(self):
        """(Re)flatten when the module was moved / re-created since the last call."""
        ok = self.flat is not None
        if ok:
            off = 0
            for p in self.params:
                if p.data.data_ptr() != self.flat.data_ptr() + off * 4 or p.data.device != self.flat.device:
                    ok = False
                    break
                off += p.numel()
        if ok:
            return False
        total = sum(p.numel() for p in self.params)
        flat = torch.zeros(total, dtype=torch.float32, device=self.ops.device)
        grad = torch.zeros(total, dtype=torch.float32, device=self.ops.device)
        off = 0
        for p in self.params:
            n = p.numel()
            flat[off:off + n].copy_(p.data.reshape(-1).to(self.ops.device, torch.float32))
            p.data = flat[off:off + n].view(p.shape)
            p.grad = grad[off:off + n].view(p.shape)
            off += n
        self.flat, self.grad = flat, grad
        # Adam moments: allocated (and zero-filled, on the stream current NOW) together with the parameters. They used to
        # be created lazily inside the first adam() call -- which for the generator runs on the update side stream,
        # while torch's zero-fill went to the main stream: the first update could read the moments before they were
        # zeroed (harmless on fresh device memory, NaN / a wrong first step on recycled blocks).
        if self.m is None or self.m.numel() != total or self.m.device != flat.device:
            self.m = torch.zeros_like(flat)
            self.v = torch.zeros_like(flat)
        return True

    def zero_grad(self):
        self.ops.memset0(self.grad)

    def adam(self, lr, beta1, beta2=0.999, eps=1e-8, grad_scale=1.0):
        lr_fn = lr if callable(lr) else (lambda: lr)

        def bump():
            self.step += 1

        self.ops.host(bump)
        self.ops.adam(self.flat, self.grad, self.m, self.v, lr_fn(), beta1, beta2, eps, self.step, grad_scale,
                      dyn=lambda: (lr_fn(), self.step))

    def versions(self):
        return tuple(p._version for p in self.params) + (self.step, self.flat.data_ptr())


class ConvL:
    """One convolution: geometry, buffers, packed weights, plans."""

    def __init__(self, eng, name, geom, Cin, Cout, weight, transposed=False, bias=None, act=0, out_f32=False,
                 need_dx=True, x_buf=None, raw_buf=None):
        ops = eng.ops
        self.eng, self.name, self.g = eng, name, geom
        self.Cin, self.Cout = Cin, Cout
        self.Cin_p, self.Cout_p = geom.in_lay.ld, geom.out_lay.ld
        self.weight, self.transposed, self.bias, self.act, self.out_f32 = weight, transposed, bias, act, out_f32
        self.need_dx = need_dx
        self.T = geom.k * geom.k
        self.x = x_buf if x_buf is not None else ops.zeros(geom.in_lay.rows, self.Cin_p)
        self.raw = raw_buf if raw_buf is not None else ops.zeros(
            geom.out_lay.rows, self.Cout_p, dtype=torch.float32 if out_f32 else None)
        self.bias_p = None
        if bias is not None:
            self.bias_p = ops.zeros(self.Cout_p, dtype=torch.float32)
        self.wp = ops.zeros(self.T, self.Cout_p, self.Cin_p)
        self.fwd = convops.fwd_plans(ops.lib, geom, self.x, self.wp, self.raw, self.Cin_p, self.Cout_p,
                                     bias=self.bias_p, act=act, out_f32=out_f32)
        self.fwd_stats = None
        self.dgrad_fused = None
        self.bwd_ready = False
        self.has_wgrad = False
        self.own_dy = False
        self.dw = None

    # weight strides (n = out channel, c = in channel, t = tap) of the fp32 master tensor
    def _strides(self):
        w = self.weight
        if self.transposed:            # [Cin][Cout][kh][kw]
            return w.stride(1), w.stride(0), 1
        return w.stride(0), w.stride(1), 1

    def pack(self, with_dgrad):
        """Immediate (one launch per operand) packing; the engines batch the same work through pack_jobs()."""
        ops = self.eng.ops
        sn, sc, st = self._strides()
        ops.pack_weight(self.weight, sn, sc, st, self.Cout, self.Cin, self.T, self.wp, self.Cout_p, self.Cin_p)
        if with_dgrad and self.need_dx and self.bwd_ready:
            ops.pack_weight(self.weight, sc, sn, st, self.Cin, self.Cout, self.T, self.wd, self.Cin_p, self.Cout_p)
        self.pack_bias()

    def pack_bias(self):
        if self.bias is not None:
            self.eng.ops.unpack_wgrad(self.bias, self.bias_p, 1, 0, 0, self.Cout, 1, 1, False)      # fp32 copy

    def pack_jobs(self):
        """Batched-kernel job descriptions (mmh_param_jobs kind 0) of this conv's tensor-core operands."""
        sn, sc, _ = self._strides()
        jobs = [dict(kind=0, src=self.weight, dst=self.wp, s_n=sn, s_c=sc, N=self.Cout, C=self.Cin, T=self.T,
                     Np=self.Cout_p, Cp=self.Cin_p)]
        if self.need_dx and self.bwd_ready:
            jobs.append(dict(kind=0, src=self.weight, dst=self.wd, s_n=sc, s_c=sn, N=self.Cin, C=self.Cout, T=self.T,
                             Np=self.Cin_p, Cp=self.Cout_p))
        return jobs

    def unpack_job(self):
        sn, sc, _ = self._strides()
        return dict(kind=1, src=self.dw, dst=self.weight.grad, s_n=sn, s_c=sc, N=self.Cout, C=self.Cin, T=self.T)

    def prepare_backward(self, dy_key, dx_key, need_wgrad=True):
        if self.bwd_ready:
            return
        ops, g = self.eng.ops, self.g
        # with side-stream weight gradients dy is read asynchronously: every conv owns its dy (no sharing)
        self.own_dy = need_wgrad and ops.side_stream is not None
        if self.own_dy:
            dy_key = (dy_key, self.name)
        self.dy = self.eng.scratch(("dy", dy_key, g.out_lay.rows, self.Cout_p), g.out_lay.rows, self.Cout_p)
        if self.need_dx:
            self.dx = self.eng.scratch(("dx", dx_key, g.in_lay.rows, self.Cin_p), g.in_lay.rows, self.Cin_p)
            self.wd = ops.zeros(self.T, self.Cin_p, self.Cout_p)
            self.dgrad = convops.dgrad_plans(ops.lib, g, self.dy, self.wd, self.dx, self.Cin_p, self.Cout_p)
        if need_wgrad and self.bias is not None:
            self.dbias = ops.zeros(2 * self.Cout_p, dtype=torch.float32)
        self.has_wgrad = need_wgrad          # the packed-gradient buffer is bound by the engine (bind_dw)
        self.bwd_ready = True

    def dw_numel(self):
        return self.T * self.Cout * self.Cin if self.has_wgrad else 0

    def bind_dw(self, flat, off):
        """Packed weight gradient [T][Cout][Cin] fp32 = a slice of the engine's flat buffer (one memset, one unpack
        launch per backward pass)."""
        n = self.dw_numel()
        self.dw = flat[off:off + n].view(self.T, self.Cout, self.Cin)
        self.wgrad = convops.wgrad_plans(self.eng.ops.lib, self.g, self.x, self.dy, self.dw, self.Cin_p, self.Cout_p,
                                         self.Cin, self.Cout)
        return off + n

    def run_fwd(self, stats=False):
        for p in (self.fwd_stats if stats else self.fwd):
            self.eng.ops.run_conv(p, (self.name, "fwd"))

    def attach_bn(self, bn):
        """Second set of forward plans whose epilogue accumulates the BatchNorm statistics of ``bn`` (training)."""
        if self.fwd_stats is None:
            self.fwd_stats = convops.fwd_plans(self.eng.ops.lib, self.g, self.x, self.wp, self.raw, self.Cin_p,
                                               self.Cout_p, bn_sums=bn.sums, bn_C=bn.C)
        return True

    def fuse_bn_bwd(self, producer, bn, relu, dropout):
        """Second set of data-gradient plans whose epilogue does the BatchNorm backward reduction of ``bn`` (the
        BatchNorm behind ``producer``, whose ReLU / dropout output this convolution reads through reflect padding):
        stride-1 reflect geometries with a contraction long enough to hide the epilogue (same rule as CONV_STATS)."""
        g = self.g
        if not (FUSE_BN_BWD and self.bwd_ready and self.need_dx and g.kind == 's1' and g.pad_mode == 'reflect'
                and self.eng.fused_stats() and self.Cin_p % 64 == 0 and self.T * self.Cout_p >= FUSE_BN_BWD_MIN_K
                and producer.g.out_lay.ld == self.Cin_p):
            return False
        if self.dgrad_fused is None:
            self.dgrad_fused = convops.dgrad_plans(
                self.eng.ops.lib, g, self.dy, self.wd, self.dx, self.Cin_p, self.Cout_p,
                bn_bwd=dict(x=producer.raw, xl=producer.g.out_lay, coef=bn.coef, save=bn.save, sums=bn.bsums, C=bn.C,
                            relu=relu, dropout=dropout))
        return True

    def run_bwd(self, want_wgrad=True, want_dx=True, fused_key=None):
        """Consumes self.dy. Weight gradient accumulates into weight.grad; data gradient lands in self.dx.
        fused_key: dropout key of the producer layer -> run the plans of fuse_bn_bwd (masked gradient + BN sums)."""
        ops = self.eng.ops
        do_w = want_wgrad and self.has_wgrad
        # The weight gradient is off the critical path (nothing before the optimiser reads it): it goes to the side
        # stream, where it overlaps the bandwidth-bound BN-backward kernels of the next layers; the data gradient,
        # which those kernels wait for, is enqueued first.
        overlap = do_w and ops.side_stream is not None and self.own_dy
        if overlap:
            ops.fork()
        if want_dx and self.need_dx:
            if fused_key is not None:
                for p in self.dgrad_fused:
                    ops.run_conv_key(p, fused_key, (self.name, "dgrad"))
            else:
                for p in self.dgrad:
                    ops.run_conv(p, (self.name, "dgrad"))
        if do_w:
            if overlap:
                with ops.side():
                    for p in self.wgrad:
                        ops.run_wgrad(p, (self.name, "wgrad"))
                self.eng.side_pending = True
            else:
                for p in self.wgrad:
                    ops.run_wgrad(p, (self.name, "wgrad"))
            if self.bias is not None:
                ops.memset0(self.dbias)
                ops.bn_stats(self.dy, self.g.out_lay.rows, self.Cout_p, self.Cout_p, self.dbias)
                ops.unpack_wgrad(self.dbias, self.bias.grad, 1, 0, 0, self.Cout, 1, 1, True)

    def dx_source(self):
        g = self.g
        return GradSource(self.dx, g.in_lay, g.in_pad_lo, g.in_pad_hi, g.pad_mode == 'reflect')


class BNL:
    """BatchNorm2d over a raw conv output (train: batch statistics; eval: running statistics)."""

    def __init__(self, eng, mod, Cc, chain=0):
        ops = eng.ops
        self.eng, self.mod, self.C = eng, mod, Cc
        self.ticket = eng.tickets[chain]         # "last block" counter of the stream this layer's kernels run on
        f = lambda n: ops.zeros(n, dtype=torch.float32)
        self.sums, self.coef, self.save, self.bsums, self.bsums_g, self.k = f(2 * Cc), f(2 * Cc), f(2 * Cc), f(2 * Cc), f(2 * Cc), f(2 * Cc)

    def forward(self, raw, rows, ld, count, training, in_epilogue=False):
        """in_epilogue: the convolution that produced ``raw`` already accumulated the sums (ConvL.attach_bn)."""
        ops, m = self.eng.ops, self.mod
        if training:
            eng = self.eng
            if in_epilogue:
                w = eng.peer_world()
                ops.bn_finalize_reset(w, self.sums, count * (w.size if w else 1), m.weight, m.bias, m.running_mean,
                                      m.running_var, BN_MOM, BN_EPS, self.C, self.coef, self.save)
            elif eng.fused_stats():
                # one launch: sums -> (peer exchange) -> coefficients; self.sums returns to zero
                w = eng.peer_world()
                ops.bn_stats_finalize(w, raw, rows, ld, self.C, self.sums, self.ticket, count * (w.size if w else 1),
                                      m.weight, m.bias, m.running_mean, m.running_var, BN_MOM, BN_EPS, self.coef,
                                      self.save)
            else:               # NCCL / gloo groups: statistics, all-reduce, finalise
                ops.memset0(self.sums)
                ops.bn_stats(raw, rows, ld, self.C, self.sums)
                count = eng.sync_stats(self.sums, count)
                ops.bn_finalize(self.sums, count, m.weight, m.bias, m.running_mean, m.running_var, BN_MOM, BN_EPS,
                                True, self.C, self.coef, self.save)
            ops.host(m.note_batch)
        else:
            ops.bn_finalize(None, 1.0, m.weight, m.bias, m.running_mean, m.running_var, BN_MOM, BN_EPS, False, self.C,
                            self.coef, self.save)

    def backward(self, dz, dz_f32, relu, dropout, key, x, xl, dy, yl, count, want_wgrad=True):
        ops, m = self.eng.ops, self.mod
        eng = self.eng
        dg, db = (m.weight.grad, m.bias.grad) if want_wgrad else (None, None)
        if eng.fused_stats():
            w = eng.peer_world()
            ops.bn_bwd_reduce_finalize(w, dz, dz_f32, relu, dropout, key, x, xl, self.coef, self.save, self.bsums,
                                       self.k, self.ticket, count * (w.size if w else 1), dg, db)
        else:
            ops.memset0(self.bsums)
            ops.bn_bwd_reduce(dz, dz_f32, relu, dropout, key, x, xl, self.coef, self.save, self.bsums)
            eng.bwd_finalize(self.bsums, self.bsums_g, count, self.k, dg, db, self.C)
        ops.bn_bwd_apply(dz, dz_f32, relu, dropout, key, x, xl, self.coef, self.save, self.k, dy, yl)

    def backward_fused(self, srcs, x, xl, dy, yl, count, want_wgrad=True):
        """The consumer's data-gradient epilogue already masked the gradient and accumulated self.bsums
        (ConvL.fuse_bn_bwd): finalise (k, dgamma, dbeta; + exchange) and apply."""
        ops, m, eng = self.eng.ops, self.mod, self.eng
        dg, db = (m.weight.grad, m.bias.grad) if want_wgrad else (None, None)
        w = eng.peer_world()
        ops.bn_bwd_finalize_reset(w, self.bsums, count * (w.size if w else 1), self.k, dg, db, self.C)
        ops.bn_bwd_apply((list(srcs), None), False, False, False, 0, x, xl, self.coef, self.save, self.k, dy, yl)


def param_store(ops: Ops, module) -> ParamStore:
    """The module's one ParamStore (created on first use; re-created only when the module moved to another Ops)."""
    st = module.__dict__.get("_mmh_store")
    if st is None or st.ops is not ops:
        st = ParamStore(ops, module)
        module.__dict__["_mmh_store"] = st
    return st


class EngineBase:
    def __init__(self, ops: Ops, module, B, H, W, world=None):
        self.ops, self.module, self.B, self.H, self.W = ops, module, B, H, W
        self._scratch = {}
        self.world = world         # None or an object with all_reduce(tensor) and size
        self.store = param_store(ops, module)
        # "last block" tickets of the one-launch statistics kernels: one per concurrently running layer chain
        self.tickets = [ops.zeros(1, dtype=torch.int32) for _ in range(3)]
        self.ticket = self.tickets[0]
        self.packed_version = None
        self.seed = 0

    def scratch(self, key, rows, ld, dtype=None):
        t = self._scratch.get(key)
        if t is None:
            t = self.ops.zeros(rows, ld, dtype=dtype)
            self._scratch[key] = t
        return t

    def chain(self, s):
        """Context: the launches of layer chain ``s`` (stream s of a PAT block / stem). Chain 0 stays on the launch
        stream; chains 1, 2 go to their own CUDA streams when the Ops has them (Ops.chains), so that the bandwidth-bound
        kernels of one chain run under the tensor-core kernels of another. fork_chains() / join_chains() bracket a
        group of concurrently running chains."""
        ch = self._chains()
        return self.ops.side(ch[s] if s < len(ch) else None)

    def _chains(self):
        """Side-stream index per chain. Data-parallel groups keep everything on the launch stream (MMH_PAT_STREAMS_DP=1
        overrides): the SyncBN exchanges -- sequence-numbered peer mailboxes, or torch.distributed collectives that follow
        torch's current stream -- assume one stream-ordered sequence of exchanges per rank."""
        w = self.world
        if w is not None and w.size > 1 and os.environ.get("MMH_PAT_STREAMS_DP", "0") != "1":
            return [None]
        return self.ops.chains

    def fork_chains(self, n=3):
        ch = self._chains()
        for s in range(1, min(n, len(ch))):
            self.ops.fork(ch[s])

    def join_chains(self, n=3):
        ch = self._chains()
        for s in range(1, min(n, len(ch))):
            self.ops.join(ch[s])

    def drop_key(self, layer_id, lay):
        """Dropout key of a layer whose output has layout ``lay``; data parallel: masks are those of the joint batch
        (this rank's samples are rank*B ... rank*B + B - 1 of it), see KeyRef."""
        w = self.world
        rank = getattr(w, "rank", 0) if (w is not None and w.size > 1) else 0
        return KeyRef(self.seed, layer_id, rank * self.B * lay.H * lay.W * ((lay.C + 7) // 8))

    def sync_stats(self, sums, count):
        if self.world is not None and self.world.size > 1:
            self.ops.host(lambda: self.world.all_reduce(sums))
            return count * self.world.size
        return count

    def sync_bwd_stats(self, local, glob, count):
        if self.world is not None and self.world.size > 1:
            def exchange():
                glob.copy_(local)
                self.world.all_reduce(glob)
            self.ops.host(exchange)
            return glob, count * self.world.size
        return local, count

    def fused_stats(self):
        """One-launch BN statistics (reduction + exchange + finalisation in the last block): always, except when the
        group exchanges through torch.distributed collectives (MMH_SYNCBN=nccl, gloo in the CPU tests)."""
        w = self.world
        return w is None or w.size <= 1 or getattr(w, "peer", None) is not None

    def peer_world(self):
        """The data-parallel group if its BN exchanges run through peer memory, else None."""
        w = self.world
        return w if (w is not None and w.size > 1 and getattr(w, "peer", None) is not None) else None

    def bwd_finalize(self, local, glob, count, k, dgamma, dbeta, Cc):
        """k = global (sum dz, sum dz*xhat) / global count; dgamma / dbeta accumulate the local sums."""
        w = self.peer_world()
        if w is not None:
            self.ops.bn_bwd_finalize_sync(w, local, glob, count * w.size, k, dgamma, dbeta, Cc)
        else:
            sg, count = self.sync_bwd_stats(local, glob, count)
            self.ops.bn_bwd_finalize(sg, local, count, k, dgamma, dbeta, Cc)

    def convs(self):
        raise NotImplementedError

    def prepare_training(self):
        """One-off allocation of the backward buffers / plans and the first packing of the weights. Callers that
        record launch tapes do this before recording, so that a replay holds only the per-step work."""
        self._prepare_backward()
        self.repack()

    def repack(self, force=False):
        """bf16 tensor-core operands of every weight (forward and, once backward is prepared, data-gradient form):
        one batched launch per optimiser step."""
        if self.store.ensure():
            self._pack_table = self._unpack_table = None        # parameter storage moved: rebuild the job tables
        v = self.store.versions()
        if not force and v == self.packed_version:
            return
        key = tuple(c.bwd_ready for c in self.convs())
        if getattr(self, "_pack_table", None) is None or self._pack_key != key:
            jobs = [j for c in self.convs() for j in c.pack_jobs()]
            self._pack_table, self._pack_key = self.ops.make_param_jobs(jobs), key
        self.ops.run_param_jobs(self._pack_table)
        for c in self.convs():
            c.pack_bias()
        self.packed_version = v

    def bind_wgrad(self):
        """Flat packed-gradient buffer behind every conv's dw + the batched unpack table."""
        convs = [c for c in self.convs() if c.has_wgrad]
        self.dw_all = self.ops.zeros(sum(c.dw_numel() for c in convs), dtype=torch.float32)
        off = 0
        for c in convs:
            off = c.bind_dw(self.dw_all, off)
        self._unpack_table = None

    def begin_wgrad(self):
        self.ops.memset0(self.dw_all)

    def bucketed(self):
        """Weight gradients of this engine are summed over the ranks bucket by bucket during backward."""
        w = self.world
        return bool(GRAD_BUCKETS and getattr(self, "buckets", None) and w is not None and w.size > 1
                    and hasattr(w, "all_reduce_bucket"))

    def reduce_bucket(self, name):
        """The packed weight gradients of bucket ``name`` are complete (their kernels are enqueued): sum over ranks."""
        if not self.bucketed():
            return
        lo, hi = self.buckets[name]
        view, w, ops = self.dw_all[lo:hi], self.world, self.ops
        ops.host(lambda: w.all_reduce_bucket(ops, view))

    def end_wgrad(self):
        if getattr(self, "side_pending", False):
            self.ops.join()            # weight gradients launched on the side stream are complete past this point
            self.side_pending = False
        if self.bucketed():
            w, ops = self.world, self.ops
            ops.host(lambda: w.wait_buckets(ops))
        if getattr(self, "_unpack_table", None) is None:
            self.store.ensure()
            self._unpack_table = self.ops.make_param_jobs([c.unpack_job() for c in self.convs() if c.has_wgrad])
        self.ops.run_param_jobs(self._unpack_table)

    def epilogue_stats(self, conv: ConvL, bn: BNL, training):
        """BatchNorm statistics inside the convolution's epilogue? (training, one-launch statistics available, bias-free
        bf16 convolution; see CONV_STATS.)"""
        if not (training and CONV_STATS and self.fused_stats() and conv.bias is None and not conv.out_f32):
            return False
        if CONV_STATS == 2 and (conv.T * conv.Cin_p < 2304 or conv.g.kind == 'up'):
            return False
        return conv.attach_bn(bn)

    def _stage_fwd(self, conv: ConvL, bn: BNL, training):
        fused = self.epilogue_stats(conv, bn, training)
        conv.run_fwd(stats=fused)
        ol = conv.g.out_lay
        bn.forward(conv.raw, ol.rows, ol.ld, self.B * ol.H * ol.W, training, in_epilogue=fused)

    def _stage_bwd(self, conv: ConvL, bn: BNL, srcs, relu, dropout, key, trunk=None, want_wgrad=True, want_dx=True,
                   dz_out_f32=None, fused=False):
        """Backward of conv -> BN -> [ReLU] -> [dropout]: gather dz from the consumers, BN backward, conv backward.
        fused: the (single) consumer's data gradient ran with ConvL.fuse_bn_bwd plans."""
        ops = self.ops
        ol = conv.g.out_lay
        B, H, W, Cc = ol.B, ol.H, ol.W, ol.C
        if fused:
            bn.backward_fused(srcs, conv.raw, ol, conv.dy, ol, B * H * W, want_wgrad)
            conv.run_bwd(want_wgrad, want_dx)
            return
        if dz_out_f32 is None and len(srcs) <= 2 and FUSE_GATHER:
            dz, f32 = (list(srcs), trunk), False          # gathered inside the two BN backward kernels
        else:
            if dz_out_f32 is not None:
                dz, f32 = dz_out_f32, True
            else:
                dz, f32 = self.scratch(("dz", conv.name, B * H * W, Cc), B * H * W, Cc), False
            ops.grad_gather(srcs, B, H, W, Cc, dz, plain_lay(B, H, W, Cc), f32, trunk=trunk)
        bn.backward(dz, f32, relu, dropout, key, conv.raw, ol, conv.dy, ol, B * H * W, want_wgrad)
        conv.run_bwd(want_wgrad, want_dx)


# ====================================================================================================
class GeneratorEngine(EngineBase):
    def __init__(self, ops, module, B, H, W, world=None):
        super().__init__(ops, module, B, H, W, world)
        m = module.model
        assert m.n_downsampling == 2, "only n_downsampling=2 is built (the shipped configuration)"
        assert H % 4 == 0 and W % 4 == 0
        ngf, self.nb, self.use_dropout = m.ngf, m.n_blocks, m.use_dropout
        assert ngf % 16 == 0, "ngf must be a multiple of 16"
        # 3 streams: MM-HAND's generator (image, pose, depth; models/Generator.py); 2 streams: the pose-transfer baseline
        # of the benchmark harness (image, pose; baselines/quantitative_on_benchmarks/networks/model_variants.py)
        self.ns = ns = getattr(m, "n_streams", 3)
        self.in_nc = [m.input_nc_s1, m.input_nc_s2] + ([m.input_nc_s3] if ns == 3 else [])
        self.out_nc = m.output_nc
        dim = ngf * 4
        self.dim = dim
        h4, w4 = H // 4, W // 4
        E = self
        # --- stems: conv7 -> BN ReLU -> conv3 s2 -> BN ReLU -> conv3 s2 -> BN ReLU
        self.stem = []
        for s in range(ns):
            seq = getattr(m, "stream%d_down" % (s + 1))
            cin = self.in_nc[s]
            c7 = ConvL(E, "s%d.c7" % s, geom_s1(B, H, W, 7, 'reflect', chan_pad(cin), ngf), cin, ngf, seq[1].weight,
                       need_dx=False)
            d1 = ConvL(E, "s%d.d1" % s, geom_s2(B, H, W, ngf, 2 * ngf), ngf, 2 * ngf, seq[4].weight)
            d2 = ConvL(E, "s%d.d2" % s, geom_s2(B, H // 2, W // 2, 2 * ngf, dim), 2 * ngf, dim, seq[7].weight)
            self.stem.append(dict(c7=c7, d1=d1, d2=d2, bn7=BNL(E, seq[2], ngf, s), bn1=BNL(E, seq[5], 2 * ngf, s),
                                  bn2=BNL(E, seq[8], dim, s)))
        # --- PAT blocks
        j = 6 if self.use_dropout else 5
        self.blocks = []
        for i in range(self.nb):
            blk = m.att[i]
            cs = dim if i == 0 else 2 * dim
            b = dict(c1=[], c2=[], bn1=[], bn2=None)
            for s in range(ns):
                seq = getattr(blk, "conv_block_stream%d" % (s + 1))
                cin = dim if s == 0 else cs
                c1 = ConvL(E, "b%d.s%d.c1" % (i, s), geom_s1(B, h4, w4, 3, 'reflect', cin, cin), cin, cin, seq[1].weight)
                c2 = ConvL(E, "b%d.s%d.c2" % (i, s), geom_s1(B, h4, w4, 3, 'reflect', cin, dim), cin, dim, seq[j].weight)
                b["c1"].append(c1)
                b["c2"].append(c2)
                b["bn1"].append(BNL(E, seq[2], cin, s))
                if s == 0:
                    b["bn2"] = BNL(E, seq[j + 1], dim)
            self.blocks.append(b)
        # --- up path
        up = m.stream1_up
        self.up1 = ConvL(E, "up1", geom_up(B, h4, w4, dim, dim // 2), dim, dim // 2, up[0].weight, transposed=True)
        self.up2 = ConvL(E, "up2", geom_up(B, 2 * h4, 2 * w4, dim // 2, ngf), dim // 2, ngf, up[3].weight,
                         transposed=True)
        self.bnu1, self.bnu2 = BNL(E, up[1], dim // 2), BNL(E, up[4], ngf)
        self.cout = ConvL(E, "out", geom_s1(B, H, W, 7, 'reflect', ngf, chan_pad(self.out_nc)), ngf, self.out_nc,
                          up[7].weight, bias=up[7].bias, act=2, out_f32=True)
        self.trunk = [ops.zeros(B * h4 * w4, dim, dtype=torch.float32) for _ in range(2)]
        self.fake = ops.zeros(B, self.out_nc, H, W, dtype=torch.float32)
        self.h4, self.w4 = h4, w4
        self.bwd_ready = False

    def convs(self):
        out = []
        for st in self.stem:
            out += [st["c7"], st["d1"], st["d2"]]
        for b in self.blocks:
            out += b["c1"] + b["c2"]
        return out + [self.up1, self.up2, self.cout]

    def _prepare_backward(self):
        if self.bwd_ready:
            return
        for s, st in enumerate(self.stem):
            st["c7"].prepare_backward("stem7.s%d" % s, None)          # per stem: the three run concurrently
            st["d1"].prepare_backward("stemd1.s%d" % s, "stemd1.s%d" % s)
            st["d2"].prepare_backward("stemd2.s%d" % s, "stemd2.s%d" % s)
        for i, b in enumerate(self.blocks):
            for s in range(self.ns):
                b["c1"][s].prepare_backward("c1.s%d" % s, "c1.s%d" % s)   # dx kept until the previous block's gate backward
                b["c2"][s].prepare_backward("c2.s%d" % s, "c2.s%d" % s)
        self.up1.prepare_backward("up1", "up1")
        self.up2.prepare_backward("up2", "up2")
        self.cout.prepare_backward("out", "out")
        self.dtrunk = self.ops.zeros(self.B * self.h4 * self.w4, self.dim, dtype=torch.float32)
        self.bind_wgrad()
        # gradient buckets in the order backward completes them: [up path + last block], each earlier block, and
        # [first block + stems] -- ranges of the packed buffer (bind order = convs(): stems, blocks, up path)
        off, start = 0, {}
        for c in self.convs():
            start[c.name] = off
            off += c.dw_numel()
        first = lambda i: start[self.blocks[i]["c1"][0].name]
        self.buckets = {}
        nb = self.nb
        if nb >= 2:
            self.buckets["tail"] = (first(nb - 1), off)
            for i in range(nb - 2, 0, -1):
                self.buckets["b%d" % i] = (first(i), first(i + 1))
            self.buckets["head"] = (0, first(1))
        else:
            self.buckets["head"] = (0, off)
        self.bwd_ready = True
        self.repack(force=True)

    # ------------------------------------------------------------------------------------------ taped inference
    def forward_taped(self, x1, x2, x3):
        """Eval-mode forward through a recorded launch sequence (the aug.py / test() path): the ~300 launches of a
        forward are recorded once per (input shapes, launch stream) and replayed -- a few microseconds of host time
        per launch instead of the argument marshalling of an eager call, which made inference host-bound on slow hosts
        (bench.py --workload infer e2e: 42 vs 20 ms per batch of 32 on two boxes of the same pool). The inputs' device
        pointers are the only per-call arguments (the three stems' assemble launches): they are patched per replay.
        Weights are re-packed before the replay when they changed (``repack`` compares parameter versions); eval-mode
        BatchNorm reads the running statistics through their (stable) pointers inside the recorded launches."""
        ops = self.ops
        if ops.tape is not None:                     # inside somebody else's recording: plain launches join that tape
            return self.forward(x1, x2, None, x3, None, False)
        ins = (x1, x2, x3)
        key = (tuple(tuple(t.shape) for t in ins), ops.st().value)
        if getattr(self, "_infer_tape", None) is None or self._infer_key != key:
            self.repack()
            with ops.record() as tape:
                out = self.forward(x1, x2, None, x3, None, False)
            # where the inputs' pointers sit in the recorded argument lists
            ptrs = {t.data_ptr(): i for i, t in enumerate(ins)}
            assert len(ptrs) == len(ins), "the generator's three inputs must be distinct tensors"
            slots = []
            for ci, (fn, args, _) in enumerate(tape.cmds):
                if fn is None:
                    continue
                for ai, v in enumerate(args):          # (pointers travel as plain integers, kernels._p)
                    if isinstance(v, int) and not isinstance(v, bool) and v in ptrs:
                        slots.append((ci, ai, ptrs[v]))
            assert len({i for _, _, i in slots}) == len(ins), "an input is not consumed by a recorded launch"
            self._infer_tape, self._infer_out, self._infer_key, self._infer_slots = tape, out, key, slots
            return out
        self.repack()
        cmds = self._infer_tape.cmds
        for ci, ai, i in self._infer_slots:
            cmds[ci][1][ai] = ins[i].data_ptr()
        # (no need to keep `ins` alive past this call: the launches that read them are ordered before anything the
        # caller enqueues later on the launch stream -- the chain streams join it inside the recorded sequence)
        self._infer_tape.replay(0)
        return self._infer_out

    # ------------------------------------------------------------------------------------------ forward
    def forward(self, x1, x2a, x2b, x3a, x3b, training, step=0, net_id=0):
        """x1: image [B,3,H,W]; (x2a | x2b): pose maps; (x3a | x3b): depth maps (NCHW fp32, second halves may be
        None when the caller already concatenated). Returns the fp32 NCHW image buffer (owned by the engine)."""
        ops, B, H, W = self.ops, self.B, self.H, self.W
        if training:
            self._prepare_backward()
        self.repack()
        self.training, self.step, self.net_id = training, step, net_id
        self.ops.step = step
        b0 = self.blocks[0]
        # the three stems, and the three streams of every PAT block, are independent layer chains: each runs on its own
        # CUDA stream (chain()), so that the bandwidth-bound kernels of one chain execute under the convolutions of another
        ns = self.ns
        self.fork_chains(ns)
        for s, (a, b_) in ((0, (x1, None)), (1, (x2a, x2b)), (2, (x3a, x3b)))[:ns]:
            with self.chain(s):
                st = self.stem[s]
                c7, d1, d2 = st["c7"], st["d1"], st["d2"]
                ops.assemble(a, b_, c7.x, c7.g.in_lay, 3, 3, True)
                self._stage_fwd(c7, st["bn7"], training)
                ops.norm_act(c7.raw, c7.g.out_lay, st["bn7"].coef, True, False, 0, d1.x, d1.g.in_lay, 1, 1, False)
                self._stage_fwd(d1, st["bn1"], training)
                ops.norm_act(d1.raw, d1.g.out_lay, st["bn1"].coef, True, False, 0, d2.x, d2.g.in_lay, 1, 1, False)
                self._stage_fwd(d2, st["bn2"], training)
                nxt = b0["c1"][s]
                ops.norm_act(d2.raw, d2.g.out_lay, st["bn2"].coef, True, False, 0, nxt.x, nxt.g.in_lay, 1, 1, True,
                             dst_f32=self.trunk[0] if s == 0 else None)
        cur = 0
        for i, b in enumerate(self.blocks):
            if i > 0:
                self.fork_chains(ns)        # (block 0 continues the stems' chains)
            for s in range(ns):
                with self.chain(s):
                    c1, c2, bn1 = b["c1"][s], b["c2"][s], b["bn1"][s]
                    self._stage_fwd(c1, bn1, training)
                    drop = training and self.use_dropout
                    key = self.drop_key(net_id * 1000 + ns * i + s, c1.g.out_lay) if drop else 0
                    ops.norm_act(c1.raw, c1.g.out_lay, bn1.coef, True, drop, key, c2.x, c2.g.in_lay, 1, 1, True)
                    fused2 = s == 0 and self.epilogue_stats(c2, b["bn2"], training)
                    c2.run_fwd(stats=fused2)
                    if s == 0:
                        bn2_fused = fused2
            self.join_chains(ns)
            c2s = b["c2"]
            ol = c2s[0].g.out_lay
            b["bn2"].forward(c2s[0].raw, ol.rows, ol.ld, B * ol.H * ol.W, training, in_epilogue=bn2_fused)
            d2b = d3b = d2l = d3l = None
            if i + 1 < self.nb:
                n = self.blocks[i + 1]["c1"]
                d1b, d1l = n[0].x, n[0].g.in_lay
                if ns == 3:     # swapped wiring (Generator.py:130 vs :278): [x3o | out] feeds stream 2, [x2o | out] stream 3
                    d2b, d2l = n[1].x, n[1].g.in_lay
                    d3b, d3l = n[2].x, n[2].g.in_lay
                else:           # two streams: [x2o | out] feeds stream 2 (model_variants.py:66-68)
                    d3b, d3l = n[1].x, n[1].g.in_lay
                lo, hi, refl = 1, 1, True
            else:
                d1b, d1l = self.up1.x, self.up1.g.in_lay
                lo, hi, refl = 0, 1, False
            ops.gate_fwd(c2s[0].raw, c2s[1].raw, c2s[2].raw if ns == 3 else None, ol, b["bn2"].coef, self.trunk[cur],
                         self.trunk[1 - cur], d1b, d1l, d2b, d2l, d3b, d3l, lo, hi, refl)
            cur = 1 - cur
        self._stage_fwd(self.up1, self.bnu1, training)
        ops.norm_act(self.up1.raw, self.up1.g.out_lay, self.bnu1.coef, True, False, 0, self.up2.x, self.up2.g.in_lay,
                     0, 1, False)
        self._stage_fwd(self.up2, self.bnu2, training)
        ops.norm_act(self.up2.raw, self.up2.g.out_lay, self.bnu2.coef, True, False, 0, self.cout.x,
                     self.cout.g.in_lay, 3, 3, True)
        self.cout.run_fwd()
        ops.grid_to_nchw(self.cout.raw, self.cout.g.out_lay, self.fake, self.out_nc)
        return self.fake

    # ------------------------------------------------------------------------------------------ backward
    def backward(self, dfake):
        """dfake: fp32 NCHW gradient of the loss w.r.t. the generated image. Accumulates parameter gradients."""
        assert self.training and self.bwd_ready
        ops, B = self.ops, self.B
        h4, w4, dim = self.h4, self.w4, self.dim
        step, net_id = self.step, self.net_id
        ops.step = step
        self.begin_wgrad()
        ops.tanh_bwd(dfake, self.fake, self.cout.dy, self.cout.g.out_lay, self.out_nc)
        self.cout.run_bwd()
        self._stage_bwd(self.up2, self.bnu2, [self.cout.dx_source()], True, False, 0)
        self._stage_bwd(self.up1, self.bnu1, [self.up2.dx_source()], True, False, 0)
        ops.grad_gather([self.up1.dx_source()], B, h4, w4, dim, self.dtrunk, plain_lay(B, h4, w4, dim), True)
        for i in range(self.nb - 1, -1, -1):
            b = self.blocks[i]
            ex2 = ex3 = None
            ns = self.ns
            if i + 1 < self.nb:
                n = self.blocks[i + 1]["c1"]
                srcs = [n[s].dx_source() for s in range(ns)]
                ops.grad_gather([srcs[0]] + [t.view(dim, dim) for t in srcs[1:]], B, h4, w4, dim, self.dtrunk,
                                plain_lay(B, h4, w4, dim), True, trunk=self.dtrunk)
                if ns == 3:     # swapped wiring: x2o of this block fed stream3 of the next one, x3o fed stream2
                    ex2, ex3 = srcs[2].view(0, dim), srcs[1].view(0, dim)
                else:
                    ex2 = srcs[1].view(0, dim)
            c2s, bn2 = b["c2"], b["bn2"]
            x3o, dy3 = (c2s[2].raw, c2s[2].dy) if ns == 3 else (None, None)
            ol = c2s[0].g.out_lay
            if self.fused_stats():
                w = self.peer_world()
                ops.gate_bwd_reduce_finalize(w, self.dtrunk, c2s[0].raw, c2s[1].raw, x3o, ol, bn2.coef, bn2.save,
                                             bn2.bsums, bn2.k, self.ticket, B * h4 * w4 * (w.size if w else 1),
                                             bn2.mod.weight.grad, bn2.mod.bias.grad)
            else:
                ops.memset0(bn2.bsums)
                ops.gate_bwd_reduce(self.dtrunk, c2s[0].raw, c2s[1].raw, x3o, ol, bn2.coef, bn2.save, bn2.bsums)
                self.bwd_finalize(bn2.bsums, bn2.bsums_g, B * h4 * w4, bn2.k, bn2.mod.weight.grad, bn2.mod.bias.grad,
                                  dim)
            ops.gate_bwd_apply(self.dtrunk, c2s[0].raw, c2s[1].raw, x3o, ol, bn2.coef, bn2.save, bn2.k, ex2, ex3,
                               c2s[0].dy, c2s[1].dy, dy3, ol)
            self.fork_chains(ns)
            for s in range(ns):
                with self.chain(s):
                    c1, c2, bn1 = b["c1"][s], b["c2"][s], b["bn1"][s]
                    drop = self.use_dropout
                    key = self.drop_key(net_id * 1000 + ns * i + s, c1.g.out_lay) if drop else 0
                    fused = c2.fuse_bn_bwd(c1, bn1, True, drop)
                    c2.run_bwd(fused_key=key if fused else None)
                    self._stage_bwd(c1, bn1, [c2.dx_source()], True, drop, key, fused=fused)
            if i >= 1:
                self.join_chains(ns)        # (block 0's chains run on into the stems)
            if i >= 1 and self.nb >= 2:
                self.reduce_bucket("tail" if i == self.nb - 1 else "b%d" % i)
        b0 = self.blocks[0]["c1"]
        for s in range(self.ns):
            with self.chain(s):
                st = self.stem[s]
                self._stage_bwd(st["d2"], st["bn2"], [b0[s].dx_source()], True, False, 0,
                                trunk=self.dtrunk if s == 0 else None)
                self._stage_bwd(st["d1"], st["bn1"], [st["d2"].dx_source()], True, False, 0)
                self._stage_bwd(st["c7"], st["bn7"], [st["d1"].dx_source()], True, False, 0, want_dx=False)
        self.join_chains(self.ns)
        self.reduce_bucket("head")
        self.end_wgrad()


# ====================================================================================================
class DiscriminatorEngine(EngineBase):
    def __init__(self, ops, module, B, H, W, world=None):
        super().__init__(ops, module, B, H, W, world)
        m = module
        nd = m.n_downsampling
        assert 0 <= nd <= 3, "n_downsampling in 0..3 (reference Discriminator.py:86-133)"
        ndf, self.nb, self.use_dropout = m.ngf, m.n_blocks, m.use_dropout
        assert ndf % 16 == 0 and H % (1 << nd) == 0 and W % (1 << nd) == 0
        self.in_nc = m.input_nc
        # channels of the stride-2 stages (reference :86-133): ndf -> 2 ndf -> 4 ndf [-> 4 ndf when n_downsampling = 3]
        chans = [ndf * (2 ** i) for i in range(min(nd, 2) + 1)] + ([ndf * 4] if nd == 3 else [])
        dim = chans[-1]
        self.dim, self.h4, self.w4 = dim, H >> nd, W >> nd      # (h4, w4: resolution of the residual trunk)
        h4, w4 = self.h4, self.w4
        E, seq = self, m.model
        self.c7 = ConvL(E, "c7", geom_s1(B, H, W, 7, 'reflect', chan_pad(self.in_nc), ndf), self.in_nc, ndf, seq[1].weight)
        self.bn7 = BNL(E, seq[2], ndf)
        self.downs, self.bnd = [], []
        for i in range(nd):
            self.downs.append(ConvL(E, "d%d" % (i + 1), geom_s2(B, H >> i, W >> i, chans[i], chans[i + 1]), chans[i],
                                    chans[i + 1], seq[4 + 3 * i].weight))
            self.bnd.append(BNL(E, seq[5 + 3 * i], chans[i + 1]))
        first_block = 4 + 3 * nd
        j = 6 if self.use_dropout else 5
        self.blocks = []
        for i in range(self.nb):
            cb = seq[first_block + i].conv_block
            c1 = ConvL(E, "r%d.c1" % i, geom_s1(B, h4, w4, 3, 'reflect', dim, dim), dim, dim, cb[1].weight)
            c2 = ConvL(E, "r%d.c2" % i, geom_s1(B, h4, w4, 3, 'reflect', dim, dim), dim, dim, cb[j].weight)
            self.blocks.append(dict(c1=c1, c2=c2, bn1=BNL(E, cb[2], dim), bn2=BNL(E, cb[j + 1], dim)))
        self.trunk = [ops.zeros(B * h4 * w4, dim, dtype=torch.float32) for _ in range(2)]
        self.bwd_ready = False

    def convs(self):
        out = [self.c7] + list(self.downs)
        for b in self.blocks:
            out += [b["c1"], b["c2"]]
        return out

    def _prepare_backward(self):
        if self.bwd_ready:
            return
        self.c7.prepare_backward("c7", "c7")
        for i, d in enumerate(self.downs):
            d.prepare_backward("d%d" % (i + 1), "d%d" % (i + 1))
        for b in self.blocks:
            b["c1"].prepare_backward("c1", "c1")
            b["c2"].prepare_backward("c2", "c2")
        self.dtrunk = self.ops.zeros(self.B * self.h4 * self.w4, self.dim, dtype=torch.float32)
        self.bind_wgrad()
        self.bwd_ready = True
        self.repack(force=True)

    def forward(self, xa, xb, training, step=0, net_id=0, need_backward=True):
        """(xa | xb): NCHW fp32 input (channel concatenation). Returns the fp32 logits as a plain NHWC buffer
        [B*h4*w4, 256] (owned by the engine)."""
        ops, B = self.ops, self.B
        if training and need_backward:
            self._prepare_backward()
        self.repack()
        self.training, self.step, self.net_id = training, step, net_id
        self.ops.step = step
        c7 = self.c7
        ops.assemble(xa, xb, c7.x, c7.g.in_lay, 3, 3, True)
        # c7 -> stride-2 stages -> first residual block: every stage normalises into the next convolution's input grid
        # (zero halo + parity planes for a stride-2 consumer, reflect halo for a block); the last one also writes the
        # fp32 trunk
        stages = [(c7, self.bn7)] + list(zip(self.downs, self.bnd))
        assert self.nb >= 1, "a discriminator without residual blocks has no consumer for its stem"
        consumers = list(self.downs) + [self.blocks[0]["c1"]]
        for k, (cv, bn) in enumerate(stages):
            self._stage_fwd(cv, bn, training)
            nxt, last = consumers[k], k == len(stages) - 1
            ops.norm_act(cv.raw, cv.g.out_lay, bn.coef, True, False, 0, nxt.x, nxt.g.in_lay, 1, 1, last,
                         dst_f32=self.trunk[0] if last else None)
        cur = 0
        for i, b in enumerate(self.blocks):
            c1, c2 = b["c1"], b["c2"]
            self._stage_fwd(c1, b["bn1"], training)
            drop = training and self.use_dropout
            key = self.drop_key(net_id * 1000 + i, c1.g.out_lay) if drop else 0
            ops.norm_act(c1.raw, c1.g.out_lay, b["bn1"].coef, True, drop, key, c2.x, c2.g.in_lay, 1, 1, True)
            self._stage_fwd(c2, b["bn2"], training)
            if i + 1 < self.nb:
                nx = self.blocks[i + 1]["c1"]
                dst, dl = nx.x, nx.g.in_lay
            else:
                dst, dl = None, None
            ops.norm_act(c2.raw, c2.g.out_lay, b["bn2"].coef, False, False, 0, dst, dl, 1, 1, True,
                         resid=self.trunk[cur], dst_f32=self.trunk[1 - cur])
            cur = 1 - cur
        self.logits = self.trunk[cur]
        return self.logits

    def backward(self, dlogits, want_wgrad=True, want_input_grad=False):
        """dlogits: fp32 plain [B*h4*w4, dim]. Returns the GradSource of the (padded) network input when asked."""
        assert self.training and self.bwd_ready
        ops, B, h4, w4, dim = self.ops, self.B, self.h4, self.w4, self.dim
        step, net_id = self.step, self.net_id
        ops.step = step
        if want_wgrad:
            self.begin_wgrad()
        dcur = dlogits
        for i in range(self.nb - 1, -1, -1):
            b = self.blocks[i]
            c1, c2 = b["c1"], b["c2"]
            ol = c2.g.out_lay
            b["bn2"].backward(dcur, True, False, False, 0, c2.raw, ol, c2.dy, ol, B * h4 * w4, want_wgrad)
            drop = self.use_dropout
            key = self.drop_key(net_id * 1000 + i, c1.g.out_lay) if drop else 0
            fused = c2.fuse_bn_bwd(c1, b["bn1"], True, drop)
            c2.run_bwd(want_wgrad, fused_key=key if fused else None)
            self._stage_bwd(c1, b["bn1"], [c2.dx_source()], True, drop, key, want_wgrad=want_wgrad, fused=fused)
            # d x_k = d x_{k+1} + fold(d pad(x_k))
            ops.grad_gather([c1.dx_source()], B, h4, w4, dim, self.dtrunk, plain_lay(B, h4, w4, dim), True, trunk=dcur)
            dcur = self.dtrunk
        stages = [(self.c7, self.bn7)] + list(zip(self.downs, self.bnd))
        for k in range(len(stages) - 1, -1, -1):
            cv, bn = stages[k]
            last = k == len(stages) - 1          # its gradient is the fp32 trunk gradient; the others gather their consumer's
            self._stage_bwd(cv, bn, [] if last else [stages[k + 1][0].dx_source()], True, False, 0,
                            trunk=dcur if last else None, want_wgrad=want_wgrad,
                            want_dx=want_input_grad if k == 0 else True)
        if want_wgrad:
            self.end_wgrad()
        return self.c7.dx_source() if want_input_grad else None


# ====================================================================================================
VGG_MEAN = (0.485, 0.456, 0.406)
VGG_STD = (0.229, 0.224, 0.225)


class VggEngine:
    """VGG19.features[0 : layers + 1] on two inputs (generated image: slot 0, target: slot 1) + perceptual loss backward.
    layers = 3 (the shipped value): conv, ReLU, conv, ReLU; 2: without the last ReLU; 1: conv, ReLU; 0: conv."""

    def __init__(self, ops: Ops, w1, b1, w2, b2, B, H, W, layers=3):
        assert 0 <= layers <= 3
        self.ops, self.B, self.H, self.W, self.layers = ops, B, H, W, layers
        self._scratch = {}
        E = self
        self.two = layers >= 2                      # second convolution present
        self.linear = layers in (0, 2)              # the features end on a convolution, not on a ReLU
        self.w = [t.detach().to(ops.device, torch.float32).contiguous() for t in ((w1, b1, w2, b2) if self.two else (w1, b1))]
        g1 = geom_s1(B, H, W, 3, 'zero', 16, 64)
        g2 = geom_s1(B, H, W, 3, 'zero', 64, 64)
        self.c1, self.c2 = [], []
        for slot in range(2):
            if self.two:
                c2 = ConvL(E, "vgg2.%d" % slot, g2, 64, 64, self.w[2], bias=self.w[3], act=0 if self.linear else 1)
                c1 = ConvL(E, "vgg1.%d" % slot, g1, 3, 64, self.w[0], bias=self.w[1], act=1, raw_buf=c2.x)
                # conv1 writes straight into conv2's zero-haloed input
                c1.fwd = convops.fwd_plans(ops.lib, g1, c1.x, c1.wp, c2.x, 16, 64, bias=c1.bias_p, act=1,
                                           out_map=(g2.in_lay.Hg * g2.in_lay.Wg, g2.in_lay.Wg, 1, 1, 1, 1),
                                           zero_invalid=False)
                self.c2.append(c2)
            else:
                c1 = ConvL(E, "vgg1.%d" % slot, g1, 3, 64, self.w[0], bias=self.w[1], act=0 if self.linear else 1)
            self.c1.append(c1)
        for c in self.c1 + self.c2:
            c.pack(False)
        mean = torch.tensor(VGG_MEAN, dtype=torch.float32)
        std = torch.tensor(VGG_STD, dtype=torch.float32)
        self.scale = (0.5 / std).to(ops.device)
        self.shift = ((0.5 - mean) / std).to(ops.device)
        self.bwd_ready = False

    def scratch(self, key, rows, ld, dtype=None):
        t = self._scratch.get(key)
        if t is None:
            t = self.ops.zeros(rows, ld, dtype=dtype)
            self._scratch[key] = t
        return t

    def prepare_training(self):
        """Backward buffers, plans and data-gradient operands of the generated-image slot (once)."""
        if self.bwd_ready:
            return
        c1 = self.c1[0]
        if self.two:
            c2 = self.c2[0]
            c2.need_dx = True
            c2.prepare_backward("v2", "v2", need_wgrad=False)
        c1.prepare_backward("v1", "v1", need_wgrad=False)
        if self.two:
            c2.pack(True)
        c1.pack(True)
        self.bwd_ready = True

    def features(self, x, slot):
        c1 = self.c1[slot]
        self.ops.assemble(x, None, c1.x, c1.g.in_lay, 1, 1, False, scale=self.scale, shift=self.shift)
        c1.run_fwd()
        if not self.two:
            return c1.raw
        self.c2[slot].run_fwd()
        return self.c2[slot].raw

    def loss_and_backward(self, fake, target, lambda_perc, mse, loss_acc, dfake):
        """*loss_acc += lambda * mean|f - t| ; dfake (NCHW fp32) += d loss / d fake (None: loss only)."""
        ops, B, H, W = self.ops, self.B, self.H, self.W
        ff = self.features(fake, 0)
        ft = self.features(target, 1)
        n = B * 64 * H * W
        c1 = self.c1[0]
        last = self.c2[0] if self.two else c1
        if dfake is None:
            ops.perc_loss(ff, ft, mse, lambda_perc / n, 0.0, loss_acc, None, linear=self.linear)
            return
        self.prepare_training()
        ops.perc_loss(ff, ft, mse, lambda_perc / n, lambda_perc / n, loss_acc, last.dy, linear=self.linear)
        if self.two:
            c2 = self.c2[0]
            c2.run_bwd(False)
            # through the ReLU of conv1 (its output lives in conv2's input buffer)
            ops.grad_gather([c2.dx_source()], B, H, W, 64, c1.dy, c1.g.out_lay, False, mask=c2.x, ml=c2.g.in_lay)
        c1.run_bwd(False)
        ops.input_grad_nchw(c1.dx_source(), self.scale, dfake, B, 3, H, W, True)
