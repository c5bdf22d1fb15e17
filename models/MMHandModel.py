"""Drop-in for the reference's ``models/MMHandModel.py`` (:26-384): same constructor (``opt``), attributes,
``set_input / forward / test / optimize_parameters / get_current_errors / get_current_visuals / save`` flow and
option names, with no APEX dependency. ``train.py`` runs unchanged with this package first on ``sys.path``.

One optimisation step is the reference's (:310-330): generator step (D_PB and D_PP scored with the *real* label,
L1 + VGG perceptual loss, ``pair_loss = L1 + (lambda_GAN*g_PB + lambda_GAN*g_PP)/2``), then D_PP, then D_PB (each
``DG_ratio`` times, real batch and pooled fake batch), three Adam optimisers (lr, beta1, 0.999). It is executed as an
explicit sequence of kernel launches on the engines of mmhand_b200.engine (no autograd tape): the discriminators'
weight gradients that the reference computes and discards during the generator step and the VGG
weight gradient (Q6) are simply not computed -- results are identical.

Data parallel (``--distributed``): one process per GPU, per-rank batch = batchSize // world (options/base_options.py:178
stays the caller's job), BatchNorm statistics and gradients all-reduced over NCCL (apex SyncBN + DDP in the reference).
bf16 storage with fp32 accumulation / statistics / master weights replaces AMP (``--opt_level`` is accepted and ignored;
``overflow`` is always False).
"""
import os
from collections import OrderedDict

import numpy as np
import torch

from losses.L1_plus_perceptualLoss import L1_plus_perceptualLoss
from mmhand_b200 import runtime
from util.image_pool import ImagePool

from .base_model import BaseModel
from .Discriminator import Discriminator
from .Generator import Generator
from .network_utils import GANLoss, get_norm_layer, get_scheduler, init_weights, print_network




class MMHandModel(BaseModel):
    def name(self):
        return 'MMHandModel'

    def __init__(self, opt):
        super(MMHandModel, self).__init__(opt)
        self.overflow = False
        self.device = self._device()
        input_nc = [opt.H_input_nc, opt.P_input_nc + opt.P_input_nc, opt.D_input_nc + opt.D_input_nc]
        self.netG = self.define_G(input_nc, opt.output_nc, opt.ngf, opt.norm, not opt.no_dropout, opt.init_type,
                                  self.gpu_ids, n_downsampling=opt.G_n_downsampling)
        if self.isTrain:
            self.netD_PB = self.define_D(opt.H_input_nc + opt.P_input_nc, opt.ndf, opt.n_layers_D, opt.norm,
                                         opt.no_lsgan, opt.init_type, self.gpu_ids, not opt.no_dropout_D,
                                         n_downsampling=opt.D_n_downsampling)
            self.netD_PP = self.define_D(opt.H_input_nc + opt.H_input_nc, opt.ndf, opt.n_layers_D, opt.norm,
                                         opt.no_lsgan, opt.init_type, self.gpu_ids, not opt.no_dropout_D,
                                         n_downsampling=opt.D_n_downsampling)
        if not self.isTrain or opt.continue_train:
            self.load_network()
        self.world = runtime.World() if getattr(opt, 'distributed', False) else None
        if self.world is not None and self.world.size > 1 and self.device.type == 'cuda':
            # SyncBN statistics travel through NVLink peer mailboxes inside the BN kernels (collective set-up)
            self.world.enable_peer(runtime.get_ops(self.device))
        if self.isTrain:
            self.old_lr = opt.lr
            self.fake_PP_pool = ImagePool(opt.pool_size)
            self.fake_PB_pool = ImagePool(opt.pool_size)
            self.criterionGAN = GANLoss(use_lsgan=not opt.no_lsgan, gpu=self.opt.local_rank)
            if opt.L1_type == 'origin':
                # accepted exactly as far as the reference accepts it: construction works, the first generator step
                # fails -- backward_G indexes the loss (`losses[0]`, reference :247-248) and nn.L1Loss returns a 0-dim
                # tensor, which raises IndexError in every torch release since 0.4
                self.criterionL1 = torch.nn.L1Loss()
            elif opt.L1_type == 'l1_plus_perL1':
                self.criterionL1 = L1_plus_perceptualLoss(opt.lambda_A, opt.lambda_B, opt.perceptual_layers,
                                                          self.gpu_ids, opt.percep_is_l1).to(self.device)
            else:
                raise Exception('Unsurportted type of L1!')
            # torch optimisers are kept as the holders of lr / schedulers (API compatibility); the update itself is
            # the fused Adam kernel over each network's flat parameter buffer
            mk = lambda net: torch.optim.Adam(net.parameters(), lr=opt.lr, betas=(opt.beta1, 0.999))
            self.optimizer_G, self.optimizer_D_PB, self.optimizer_D_PP = mk(self.netG), mk(self.netD_PB), mk(self.netD_PP)
            self.optimizers = [self.optimizer_G, self.optimizer_D_PB, self.optimizer_D_PP]
            self.schedulers = [get_scheduler(o, opt) for o in self.optimizers]
            self._step = 0
            self._acc = None
        self._sync_initial_state()
        if self.master:
            print('---------- Networks initialized -------------')
            print_network(self.netG)
            if self.isTrain:
                print_network(self.netD_PB)
                print_network(self.netD_PP)
                print(opt.local_rank)
            print('-----------------------------------------------')

    def _sync_initial_state(self):
        """Data parallel: every replica starts from rank 0's parameters and BatchNorm buffers, as apex
        DistributedDataParallel does at construction (reference :110-116) -- the reference never seeds its ranks, so
        identical random initialisation cannot be assumed. Includes the (frozen) VGG slice of the perceptual loss."""
        if self.world is None or self.world.size <= 1:
            return
        nets = [self.netG]
        if self.isTrain:
            nets += [self.netD_PB, self.netD_PP]
            if hasattr(self.criterionL1, 'vgg_submodel'):
                nets.append(self.criterionL1.vgg_submodel)
        with torch.no_grad():
            for net in nets:
                for t in list(net.parameters()) + list(net.buffers()):
                    self.world.dist.broadcast(t.data, 0)

    def _device(self):
        lr = self.opt.local_rank
        if isinstance(lr, str):
            return torch.device(lr)
        return torch.device('cuda', lr) if torch.cuda.is_available() else torch.device('cpu')

    def define_G(self, input_nc, output_nc, ngf, norm='batch', use_dropout=False, init_type='normal', gpu_ids=[],
                 n_downsampling=2):
        assert len(input_nc) == 3
        netG = Generator(input_nc, output_nc, ngf, norm_layer=get_norm_layer(norm_type=norm), use_dropout=use_dropout,
                         n_blocks=9, gpu_ids=gpu_ids, n_downsampling=n_downsampling)
        # N(0, 0.02) initialisation on the host, then one copy per tensor (the reference initialises after .cuda():
        # the same distribution from another generator; ~150 tiny normal_ launches less)
        init_weights(netG, init_type=init_type)
        return netG.to(self.device)

    def define_D(self, input_nc, ndf, n_layers_D=3, norm='batch', use_sigmoid=False, init_type='normal', gpu_ids=[],
                 use_dropout=False, n_downsampling=2):
        netD = Discriminator(input_nc, ndf, norm_layer=get_norm_layer(norm_type=norm), use_dropout=use_dropout,
                             n_blocks=n_layers_D, gpu_ids=[], padding_type='reflect', use_sigmoid=False,
                             n_downsampling=n_downsampling)
        init_weights(netD, init_type=init_type)
        return netD.to(self.device)

    # ------------------------------------------------------------------------------------------ data
    def set_input(self, input):
        """H2D of the six tensors (reference :200-213) straight into static device buffers, so that the recorded
        launch tapes (fixed pointers) stay valid from step to step.

        Besides the reference's fp32 tensors ('H1', 'P1', 'D1', 'H2', 'P2', 'D2': [B, C, H, W]) each input may arrive
        in the compact form of the device-side input pipeline (mmhand_b200/loader.py), in which case the
        arithmetic of the reference's dataset workers (data/generic_dataset.py:133-159) runs on the GPU, bit-identically:
          'P1_uv' / 'P2_uv'  [B, 21, 2] keypoints (x, y) -> 21 Gaussian heatmaps (mmh_heatmap_rasterize, sigma 6):
                             336 bytes per pose cross the bus instead of 5.5 MB;
          'H1_u8' / 'H2_u8'  [B, H, W, 3] uint8 frames (RGB; BGR as cv2.imread returns them with input['u8_bgr'])
                             -> ((x / 255) - 0.5) / 0.5 (mmh_image_unpack_u8);
          'D1_u8' / 'D2_u8'  [B, H, W, 3] uint8 depth frames as cv2.imread returns them -> (256*G + R) / 700 mapped to
                             [-1, 1], three equal channels (mmh_depth_unpack_u8)."""
        dev = self.device
        names = ('H1', 'P1', 'D1', 'H2', 'P2', 'D2')
        src = {}
        for k in names:
            if k in input:
                src[k] = ('f32', input[k])
            elif k[0] == 'P' and (k + '_uv') in input:
                src[k] = ('uv', input[k + '_uv'])
            elif k[0] != 'P' and (k + '_u8') in input:
                src[k] = ('u8', input[k + '_u8'])
            else:
                raise KeyError("set_input: no '%s' (nor its compact form) in the batch" % k)
        kind, t = src['H1']
        B, H, W = (t.shape[0], t.shape[2], t.shape[3]) if kind == 'f32' else (t.shape[0], t.shape[1], t.shape[2])

        def shape_of(k):
            kind, t = src[k]
            if kind == 'f32':
                return tuple(t.shape)
            return (B, t.shape[1] if kind == 'uv' else 3, H, W)

        shapes = tuple(shape_of(k) for k in names)
        if getattr(self, '_in_shapes', None) != shapes:
            self._in = {k: torch.empty(sh, dtype=torch.float32, device=dev) for k, sh in zip(names, shapes)}
            self._in_stage = {}
            self._in_shapes = shapes
            self._tapes = None
        for k in names:
            if src[k][0] == 'f32':
                self._in[k].copy_(src[k][1], non_blocking=True)
        compact = [k for k in names if src[k][0] != 'f32']
        if compact:
            ops = runtime.get_ops(dev)
            from mmhand_b200.rasterize import get_heatmaps
            for k in compact:
                kind, t = src[k]
                if kind == 'uv':
                    get_heatmaps(t, (H, W), sigma=float(input.get('sigma', 6.0)), out=self._in[k], device=dev)
                    continue
                st = self._in_stage.get(k)
                if st is None or st.shape != t.shape:
                    st = self._in_stage[k] = torch.empty(tuple(t.shape), dtype=torch.uint8, device=dev)
                st.copy_(t, non_blocking=True)
                if k[0] == 'H':
                    ops.image_unpack_u8(st, self._in[k], swap_rb=bool(input.get('u8_bgr', False)))
                else:
                    ops.depth_unpack_u8(st, self._in[k], hi=1, lo=2, div=float(input.get('depth_div', 700.0)))
        self.input_H1, self.input_P1, self.input_D1 = self._in['H1'], self._in['P1'], self._in['D1']
        self.input_H2, self.input_P2, self.input_D2 = self._in['H2'], self._in['P2'], self._in['D2']
        if 'H1_path' in input:
            self.image_paths = input['H1_path'][0] + '___' + input['H2_path'][0]

    def _g_engine(self):
        B, _, H, W = self.input_H1.shape
        return self.netG.engine(B, H, W, self.world)

    def forward(self):
        eng = self._g_engine()
        eng.seed = getattr(self.opt, 'seed', 0) if self.isTrain else 0
        self.fake_p2 = eng.forward(self.input_H1, self.input_P1, self.input_P2, self.input_D1, self.input_D2,
                                   self.netG.training, step=getattr(self, '_step', 0), net_id=0)

    def test(self):
        self.forward()

    def get_image_paths(self):
        return self.image_paths

    # ------------------------------------------------------------------------------------------ training
    def _d_engine(self, net):
        B, _, H, W = self.input_H1.shape
        eng = net.engine(B, H, W, self.world)
        eng.seed = getattr(self.opt, 'seed', 0)
        return eng

    def _lr(self, optimizer):
        return optimizer.param_groups[0]['lr']

    def _adam(self, eng, optimizer):
        """Gradient all-reduce (data parallel), fused Adam over the flat parameter buffer, repack of the bf16
        tensor-core operands. Recordable: lr and the step count are re-read on every tape replay."""
        ops = eng.ops
        scale = 1.0
        if self.world is not None and self.world.size > 1:
            stream = ops.in_side                     # NCCL follows torch's current stream
            # bucketed engines summed their convolution weight gradients over the ranks during backward: what is left is
            # the small leading region of the flat gradient (BatchNorm scales / shifts, biases)
            grad = eng.store.grad[:eng.store.n_small] if eng.bucketed() else eng.store.grad
            self.grad_sync_mode = "bucketed (%d buckets during G backward) + per-network" % len(eng.buckets) \
                if eng.bucketed() else getattr(self, "grad_sync_mode", "per-network after backward")

            def all_reduce():
                if stream is None:
                    self.world.all_reduce(grad)
                else:
                    with torch.cuda.stream(stream):
                        self.world.all_reduce(grad)
            ops.host(all_reduce)
            scale = 1.0 / self.world.size
        eng.store.adam(lambda: self._lr(optimizer), self.opt.beta1, 0.999, 1e-8, grad_scale=scale)
        eng.repack(force=True)

    def backward_G(self):
        """reference :236-261; accumulators: 0 g_PB, 1 g_PP (mean BCE), 2 lambda_A*L1, 3 lambda_B*perceptual."""
        opt, ops, acc = self.opt, self._ops, self._acc
        fake, dfake = self.fake_p2, self._dfake
        ops.memset0(dfake)
        B, C3, H, W = fake.shape
        for k, (net, other, ids) in enumerate(((self.netD_PB, self.input_P2, 1), (self.netD_PP, self.input_H1, 4))):
            eng = self._d_engine(net)
            logits = eng.forward(fake, other, True, step=self._step, net_id=ids)
            n = logits.numel()
            dl = self._dlogits(logits)
            ops.bce_logits(logits, 1.0, 1.0 / n, opt.lambda_GAN / 2.0 / n, acc[k:k + 1], dl)
            src = eng.backward(dl, want_wgrad=False, want_input_grad=True)
            ops.input_grad_nchw(src, None, dfake, B, C3, H, W, True)
        n = fake.numel()
        crit = self.criterionL1
        if opt.L1_type == 'origin':
            raise IndexError("invalid index of a 0-dim tensor. Use `tensor.item()` in Python or `tensor.item<T>()` in "
                             "C++ to convert a 0-dim tensor to a number  [L1_type='origin': the reference's backward_G "
                             "(models/MMHandModel.py:247-248) indexes the scalar nn.L1Loss returns]")
        ops.l1(fake, self.input_H2, opt.lambda_A / n, opt.lambda_A / n, acc[2:3], dfake)
        crit.vgg_engine(B, H, W).loss_and_backward(fake, self.input_H2, opt.lambda_B, opt.percep_is_l1 != 1, acc[3:4],
                                                   dfake)
        self._g_engine().backward(dfake)

    def _dlogits(self, logits):
        if getattr(self, '_dl', None) is None or self._dl.shape != logits.shape:
            self._dl = torch.empty_like(logits)
        return self._dl

    def backward_D_basic(self, netD, real_a, real_b, fake, acc_idx, ids):
        """reference :263-274: 0.5 * lambda_GAN * (BCE(D(real), 1) + BCE(D(fake), 0))."""
        opt, ops, acc = self.opt, self._ops, self._acc
        eng = self._d_engine(netD)
        for j, (xa, xb, target, nid) in enumerate(((real_a, real_b, 1.0, ids), (fake, None, 0.0, ids + 1))):
            logits = eng.forward(xa, xb, True, step=self._step, net_id=nid)
            n = logits.numel()
            dl = self._dlogits(logits)
            s = 0.5 * opt.lambda_GAN / n
            ops.bce_logits(logits, target, s, s, acc[acc_idx + j:acc_idx + j + 1], dl)
            eng.backward(dl, want_wgrad=True, want_input_grad=False)
        return eng

    def _pool_inputs(self, which=('pp', 'pb')):
        """ImagePool queries (host RNG, reference :279-289) into static buffers."""
        for w in which:
            pool, other = ((self.fake_PP_pool, self.input_H1) if w == 'pp' else (self.fake_PB_pool, self.input_P2))
            cand = torch.cat((self.fake_p2, other), 1).data
            name = '_pool_' + w
            if getattr(self, name, None) is None or getattr(self, name).shape != cand.shape:
                setattr(self, name, torch.empty_like(cand))
                self._tapes = None
            pool.query(cand, out=getattr(self, name))

    def backward_D_PB(self):
        return self.backward_D_basic(self.netD_PB, self.input_H2, self.input_P2, self._pool_pb, 6, 2)

    def backward_D_PP(self):
        return self.backward_D_basic(self.netD_PP, self.input_H2, self.input_H1, self._pool_pp, 4, 5)

    def _segment_G(self):
        ops = self._ops
        ops.memset0(self._acc)
        self.forward()
        g_eng = self._g_engine()
        g_eng.store.zero_grad()
        self.backward_G()
        dp = self.world is not None and self.world.size > 1       # (see runtime.World.tame_launches)
        if ops.side_stream is not None and os.environ.get("MMH_G_UPDATE_STREAM", "0" if dp else "1") != "0":
            # Nothing in the discriminator segments reads the generator's weights: its gradient all-reduce, Adam
            # update and operand repacking (bandwidth-bound) run on the side stream under the discriminators'
            # convolutions; _segment_D joins before the step ends.
            ops.fork(1)
            with ops.side(1):
                self._adam(g_eng, self.optimizer_G)
            self._g_update_pending = True
        else:
            self._adam(g_eng, self.optimizer_G)

    def _segment_D(self, which=('pp', 'pb')):
        ops = self._ops
        table = {'pp': (self.netD_PP, self.backward_D_PP, self.optimizer_D_PP, 4),
                 'pb': (self.netD_PB, self.backward_D_PB, self.optimizer_D_PB, 6)}
        for w in which:
            net, bwd, optim, lo = table[w]
            eng = self._d_engine(net)
            eng.store.zero_grad()
            ops.memset0(self._acc[lo:lo + 2])
            bwd()
            self._adam(eng, optim)
        if getattr(self, '_g_update_pending', False):
            ops.join(1)
            self._g_update_pending = False

    def optimize_parameters(self):
        ops = runtime.get_ops(self.device)
        hp = getattr(ops, 'main_stream', None)
        if hp is None:
            return self._optimize_parameters()
        cur = torch.cuda.current_stream(self.device)
        hp.wait_stream(cur)
        with torch.cuda.stream(hp):
            self._optimize_parameters()
        cur.wait_stream(hp)

    def _align_ranks(self):
        """Data parallel: all ranks enter the first step of a (re)built engine set together. Engine construction
        (hundreds of plans, lazy module loading on a freshly started box) can skew the ranks by many seconds; the
        in-kernel SyncBN exchange of the first layer would otherwise spin for that long (csrc/peer.cuh time-out)."""
        if self.world is not None and self.world.size > 1:
            if self.device.type == 'cuda':
                torch.cuda.synchronize(self.device)
            self.world.dist.barrier()

    def _optimize_parameters(self):
        """reference :310-330. The first call runs eagerly while recording the launch sequence of the generator
        segment and of the two discriminator segments; later calls replay the tapes (use_tape=False: always eager).
        The image-pool queries between the segments stay on the host, as in the reference."""
        self._ops = runtime.get_ops(self.device)
        ops = self._ops
        if self._acc is None:
            self._acc = torch.zeros(8, dtype=torch.float32, device=ops.device)
        B, C3, H, W = self.input_H1.shape
        if getattr(self, '_dfake', None) is None or self._dfake.shape != (B, self.opt.output_nc, H, W):
            self._dfake = torch.zeros(B, self.opt.output_nc, H, W, dtype=torch.float32, device=ops.device)
            self._tapes = None
        use_tape = getattr(self, 'use_tape', True) and self.opt.DG_ratio == 1
        tapes = getattr(self, '_tapes', None)
        # what a recorded sequence bakes in besides pointers: train / eval mode of the networks and the loss scales
        opt = self.opt
        tape_key = (self.netG.training, self.netD_PB.training, self.netD_PP.training, opt.lambda_A, opt.lambda_B,
                    opt.lambda_GAN, opt.percep_is_l1, getattr(opt, 'seed', 0))
        if tapes is not None and getattr(self, '_tape_key', None) != tape_key:
            tapes = self._tapes = None
        self._tape_key = tape_key
        if use_tape and tapes is not None and tapes[0].stream == ops._stream():
            tapes[0].replay(self._step)
            self._pool_inputs()
            tapes[1].replay(self._step)
        elif use_tape:
            # one-off work (backward buffers, plans, first weight packing) stays outside the recorded sequence
            self._g_engine().prepare_training()
            for net in (self.netD_PB, self.netD_PP):
                self._d_engine(net).prepare_training()
            if hasattr(self.criterionL1, 'vgg_engine'):
                self.criterionL1.vgg_engine(B, H, W).prepare_training()
            self._align_ranks()
            with ops.record() as tg:
                self._segment_G()
            self._pool_inputs()
            with ops.record() as td:
                self._segment_D()
            self._tapes = (tg, td)
        else:
            self._segment_G()
            for w in ('pp', 'pb'):              # reference order: D_PP DG_ratio times, then D_PB DG_ratio times
                for _ in range(self.opt.DG_ratio):
                    self._pool_inputs((w,))
                    self._segment_D((w,))
        self.overflow = False
        self._step += 1
        if self.world is not None:
            self.world.check()
        a, lam = self._acc.clone(), self.opt.lambda_GAN
        self.loss_G_GAN_PB, self.loss_G_GAN_PP = a[0], a[1]
        self.loss_originL1, self.loss_perceptual = a[2], a[3]
        self.loss_G_L1 = a[2] + a[3]
        self.pair_L1loss = self.loss_G_L1
        self.pair_GANloss = (a[0] * lam + a[1] * lam) / 2
        self.loss_D_PP = a[4] + a[5]
        self.loss_D_PB = a[6] + a[7]

    def get_current_errors(self):
        ret_errors = OrderedDict([('pair_L1loss', self.pair_L1loss)])
        ret_errors['D_PP'] = self.loss_D_PP
        ret_errors['D_PB'] = self.loss_D_PB
        ret_errors['pair_GANloss'] = self.pair_GANloss
        ret_errors['origin_L1'] = self.loss_originL1
        ret_errors['perceptual'] = self.loss_perceptual
        return ret_errors

    def get_current_visuals(self):
        import util.util as util
        height, width = self.input_H1.size(2), self.input_H1.size(3)
        panels = [util.tensor2im(self.input_H1.data), util.draw_pose_from_map(self.input_P1.data),
                  util.tensor2im(self.input_D1.data), util.tensor2im(self.input_H2.data),
                  util.draw_pose_from_map(self.input_P2.data), util.tensor2im(self.input_D2.data),
                  util.tensor2im(self.fake_p2.data)]
        vis = np.zeros((height, width * 7, 3)).astype(np.uint8)
        for i, p in enumerate(panels):
            vis[:, width * i:width * (i + 1), :] = p
        return OrderedDict([('vis', vis)])

    def save(self, label):
        self.save_network(self.netG, 'netG', label, self.gpu_ids)
        self.save_network(self.netD_PB, 'netD_PB', label, self.gpu_ids)
        self.save_network(self.netD_PP, 'netD_PP', label, self.gpu_ids)

    def pprint(self, msg):
        if self.master:
            print(msg)
