"""Drop-in for the reference's ``models/Generator.py``: same classes, constructor signatures, ``forward(input)``
contract and ``state_dict`` layout (so ``aug.py`` and old checkpoints load unchanged), with the arithmetic executed
by mmhand_b200.engine.GeneratorEngine on hand-written sm_100a kernels (NHWC bf16, fp32 accumulation).

Reference: models/Generator.py:8-130 (PATBlock), :133-283 (PATNModel), :286-313 (Generator).
"""
import functools
import os

import torch
import torch.nn as nn

from mmhand_b200 import runtime
from mmhand_b200.engine import GeneratorEngine
from mmhand_b200.modules import Conv2dParams, ConvTranspose2dParams, Slot, norm_kind, norm_params


def _use_bias(norm_layer):
    f = norm_layer.func if isinstance(norm_layer, functools.partial) else norm_layer
    return f == nn.InstanceNorm2d


def _check(norm_layer, padding_type):
    norm_kind(norm_layer)            # batch | instance (instance: accepted for checkpoint compatibility, forward raises)
    if padding_type != 'reflect':
        raise NotImplementedError('padding [%s] is not implemented' % padding_type)


class PATBlock(nn.Module):
    """Parameter layout of one pose-attention block (reference :8-113); compute lives in the engine."""

    def __init__(self, dim, padding_type, norm_layer, use_dropout, use_bias, cated_stream2=False):
        super().__init__()
        _check(norm_layer, padding_type)
        self.conv_block_stream1 = self.build_conv_block(dim, padding_type, norm_layer, use_dropout, use_bias,
                                                        cal_att=False)
        self.conv_block_stream2 = self.build_conv_block(dim, padding_type, norm_layer, use_dropout, use_bias,
                                                        cal_att=True, cated_stream2=cated_stream2)
        self.conv_block_stream3 = self.build_conv_block(dim, padding_type, norm_layer, use_dropout, use_bias,
                                                        cal_att=True, cated_stream2=cated_stream2)

    def build_conv_block(self, dim, padding_type, norm_layer, use_dropout, use_bias, cated_stream2=False,
                         cal_att=False):
        cin = dim * 2 if cated_stream2 else dim
        seq = [Slot('ReflectionPad2d(1)'), Conv2dParams(cin, cin, 3, use_bias), norm_params(norm_layer, cin), Slot('ReLU')]
        if use_dropout:
            seq.append(Slot('Dropout(0.5)'))
        seq.append(Slot('ReflectionPad2d(1)'))
        if cal_att:
            seq.append(Conv2dParams(cin, dim, 3, use_bias))
        else:
            seq += [Conv2dParams(dim, dim, 3, use_bias), norm_params(norm_layer, dim)]
        return nn.Sequential(*seq)

    def forward(self, x1, x2, x3):
        raise RuntimeError("PATBlock is executed by the fused generator engine; call Generator.forward")


class PATNModel(nn.Module):
    def __init__(self, input_nc, output_nc, ngf=64, norm_layer=nn.BatchNorm2d, use_dropout=False, n_blocks=6,
                 gpu_ids=[], padding_type='reflect', n_downsampling=2):
        assert (n_blocks >= 0 and type(input_nc) == list)
        super().__init__()
        _check(norm_layer, padding_type)
        self.input_nc_s1, self.input_nc_s2, self.input_nc_s3 = input_nc
        self.output_nc, self.ngf, self.gpu_ids = output_nc, ngf, gpu_ids
        self.n_blocks, self.use_dropout, self.n_downsampling = n_blocks, use_dropout, n_downsampling
        use_bias = _use_bias(norm_layer)

        def down(cin):
            seq = [Slot('ReflectionPad2d(3)'), Conv2dParams(cin, ngf, 7, use_bias), norm_params(norm_layer, ngf), Slot('ReLU')]
            for i in range(n_downsampling):
                mult = 2 ** i
                seq += [Conv2dParams(ngf * mult, ngf * mult * 2, 3, use_bias, stride=2),
                        norm_params(norm_layer, ngf * mult * 2), Slot('ReLU')]
            return nn.Sequential(*seq)

        self.stream1_down = down(self.input_nc_s1)
        self.stream2_down = down(self.input_nc_s2)
        self.stream3_down = down(self.input_nc_s3)
        mult = 2 ** n_downsampling
        self.att = nn.ModuleList([
            PATBlock(ngf * mult, padding_type=padding_type, norm_layer=norm_layer, use_dropout=use_dropout,
                     use_bias=use_bias, cated_stream2=(i > 0)) for i in range(n_blocks)])
        up = []
        for i in range(n_downsampling):
            mult = 2 ** (n_downsampling - i)
            up += [ConvTranspose2dParams(ngf * mult, int(ngf * mult / 2), 3, use_bias),
                   norm_params(norm_layer, int(ngf * mult / 2)), Slot('ReLU')]
        up += [Slot('ReflectionPad2d(3)'), Conv2dParams(ngf, output_nc, 7, True), Slot('Tanh')]
        self.stream1_up = nn.Sequential(*up)

    def forward(self, input):
        raise RuntimeError("PATNModel is executed by the fused generator engine; call Generator.forward")


class _GeneratorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, want_grad, anchor, x1, x2, x3):
        training = mod.training
        eng = mod.engine(x1.shape[0], x1.shape[2], x1.shape[3])
        if not training and not want_grad and os.environ.get("MMH_INFER_TAPE", "1") != "0":
            out = eng.forward_taped(x1, x2, x3)          # recorded launch sequence (aug.py / test())
        else:
            out = eng.forward(x1, x2, None, x3, None, training, step=mod._step, net_id=0)
        if training:
            mod._step += 1
        ctx.eng = eng if want_grad else None
        return out.clone()

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.eng is not None:
            ctx.eng.backward(grad_out.contiguous().float())
        return None, None, None, None, None, None


class Generator(nn.Module):
    def __init__(self, input_nc, output_nc, ngf=64, norm_layer=nn.BatchNorm2d, use_dropout=False, n_blocks=6,
                 gpu_ids=[], padding_type='reflect', n_downsampling=2):
        super().__init__()
        assert type(input_nc) == list and len(input_nc) == 3, \
            'The AttModule take input_nc in format of list only!!'
        self.gpu_ids = gpu_ids
        self.model = PATNModel(input_nc, output_nc, ngf, norm_layer, use_dropout, n_blocks, gpu_ids, padding_type,
                               n_downsampling=n_downsampling)
        self._engines = {}
        self._step = 0
        self._norm = norm_kind(norm_layer)

    def engine(self, B, H, W, world=None):
        ops = runtime.get_ops(next(self.parameters()).device)
        key = (B, H, W, str(ops.device))
        eng = self._engines.get(key)
        if eng is None:
            self._engines.clear()        # one resident engine per module: activations dominate memory
            eng = GeneratorEngine(ops, self, B, H, W, world)
            self._engines[key] = eng
        return eng

    def forward(self, input):
        if self._norm != 'batch':
            raise NotImplementedError("only norm='batch' (the shipped configuration) is computed on the B200 path; "
                                      "norm='instance' models are constructed for checkpoint compatibility only")
        x1, x2, x3 = [t.contiguous().float() for t in input]
        anchor = self.model.stream1_up[-2].bias
        want_grad = self.training and torch.is_grad_enabled()
        return _GeneratorFn.apply(self, want_grad, anchor, x1, x2, x3)
