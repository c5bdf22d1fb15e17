"""Drop-in for the reference's ``models/Discriminator.py`` (ResnetBlock :8-55, Discriminator :58-154): same
constructor, ``forward(x)`` and ``state_dict`` layout; compute on mmhand_b200.engine.DiscriminatorEngine.
The output is the last residual block's 4*ndf-channel feature map of logits (there is no 1-channel head)."""
import functools

import torch
import torch.nn as nn

from mmhand_b200 import runtime
from mmhand_b200.engine import DiscriminatorEngine
from mmhand_b200.modules import Conv2dParams, Slot, norm_kind, norm_params


class ResnetBlock(nn.Module):
    def __init__(self, dim, padding_type, norm_layer, use_dropout, use_bias):
        super().__init__()
        self.conv_block = self.build_conv_block(dim, padding_type, norm_layer, use_dropout, use_bias)

    def build_conv_block(self, dim, padding_type, norm_layer, use_dropout, use_bias):
        if padding_type != 'reflect':
            raise NotImplementedError('padding [%s] is not implemented' % padding_type)
        seq = [Slot('ReflectionPad2d(1)'), Conv2dParams(dim, dim, 3, use_bias), norm_params(norm_layer, dim), Slot('ReLU')]
        if use_dropout:
            seq.append(Slot('Dropout(0.5)'))
        seq += [Slot('ReflectionPad2d(1)'), Conv2dParams(dim, dim, 3, use_bias), norm_params(norm_layer, dim)]
        return nn.Sequential(*seq)

    def forward(self, x):
        raise RuntimeError("ResnetBlock is executed by the fused discriminator engine; call Discriminator.forward")


class _DiscriminatorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, want_grad, anchor, x):
        training = mod.training
        B, _, H, W = x.shape
        eng = mod.engine(B, H, W)
        logits = eng.forward(x, None, training, step=mod._step, net_id=mod.drop_net_id, need_backward=want_grad)
        if training:
            mod._step += 1
        ctx.eng = eng if want_grad else None
        ctx.need_in = x.requires_grad
        ctx.shape = x.shape
        return logits.view(B, eng.h4, eng.w4, eng.dim).permute(0, 3, 1, 2).contiguous()

    @staticmethod
    def backward(ctx, g):
        if ctx.eng is None:
            return None, None, None, None
        eng = ctx.eng
        B, C, H, W = ctx.shape
        dl = g.permute(0, 2, 3, 1).contiguous().float().view(-1, eng.dim)
        src = eng.backward(dl, want_wgrad=True, want_input_grad=ctx.need_in)
        gx = None
        if ctx.need_in:
            gx = torch.empty(B, C, H, W, dtype=torch.float32, device=g.device)
            eng.ops.input_grad_nchw(src, None, gx, B, C, H, W, False)
        return None, None, None, gx


class Discriminator(nn.Module):
    def __init__(self, input_nc, ngf=64, norm_layer=nn.BatchNorm2d, use_dropout=False, n_blocks=6, gpu_ids=[],
                 padding_type='reflect', use_sigmoid=False, n_downsampling=2):
        assert (n_blocks >= 0)
        super().__init__()
        self._norm = norm_kind(norm_layer)
        if use_sigmoid:
            raise NotImplementedError("use_sigmoid=True is never reached by the reference (MMHandModel.py:190)")
        self.input_nc, self.ngf, self.gpu_ids = input_nc, ngf, gpu_ids
        self.n_blocks, self.use_dropout, self.n_downsampling = n_blocks, use_dropout, n_downsampling
        f = norm_layer.func if isinstance(norm_layer, functools.partial) else norm_layer
        use_bias = f == nn.InstanceNorm2d
        seq = [Slot('ReflectionPad2d(3)'), Conv2dParams(input_nc, ngf, 7, use_bias), norm_params(norm_layer, ngf), Slot('ReLU')]
        if n_downsampling > 3:
            raise NotImplementedError("n_downsampling=%d: the reference builds 0..3 stride-2 stages (:86-133)" % n_downsampling)
        # reference :86-133: ndf -> 2 ndf -> 4 ndf, and a third stage 4 ndf -> 4 ndf when n_downsampling == 3
        chans = [ngf * (2 ** i) for i in range(min(n_downsampling, 2) + 1)] + ([ngf * 4] if n_downsampling == 3 else [])
        for i in range(n_downsampling):
            seq += [Conv2dParams(chans[i], chans[i + 1], 3, use_bias, stride=2), norm_params(norm_layer, chans[i + 1]),
                    Slot('ReLU')]
        mult = chans[-1] // ngf
        for i in range(n_blocks):
            seq.append(ResnetBlock(ngf * mult, padding_type=padding_type, norm_layer=norm_layer,
                                   use_dropout=use_dropout, use_bias=use_bias))
        self.model = nn.Sequential(*seq)
        self._engines = {}
        self._step = 0
        self.drop_net_id = 1

    def engine(self, B, H, W, world=None):
        ops = runtime.get_ops(next(self.parameters()).device)
        key = (B, H, W, str(ops.device))
        eng = self._engines.get(key)
        if eng is None:
            self._engines.clear()
            eng = DiscriminatorEngine(ops, self, B, H, W, world)
            self._engines[key] = eng
        return eng

    def forward(self, input):
        if self._norm != 'batch':
            raise NotImplementedError("only norm='batch' (the shipped configuration) is computed on the B200 path; "
                                      "norm='instance' models are constructed for checkpoint compatibility only")
        x = input.contiguous().float()
        anchor = self.model[1].weight
        want_grad = self.training and torch.is_grad_enabled()
        return _DiscriminatorFn.apply(self, want_grad, anchor, x)
