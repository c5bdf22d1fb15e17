"""Drop-in for the reference's ``models/network_utils.py``: weight initialisers, norm-layer factory, LR
schedulers, ``print_network`` and ``GANLoss`` with the reference's names and behaviour.

Reference: models/network_utils.py:12-71 (initialisers), :74-84 (get_norm_layer), :87-109 (get_scheduler),
:112-117 (print_network), :129-163 (GANLoss -- always BCE-with-logits, ``use_lsgan`` is accepted and ignored).
"""
import functools

import torch
import torch.nn as nn
from torch.nn import init
from torch.optim import lr_scheduler


def _init_by_name(m, conv_init):
    name = m.__class__.__name__
    if name.find('Conv') != -1 and getattr(m, 'weight', None) is not None:
        conv_init(m.weight.data)
    elif name.find('Linear') != -1:
        conv_init(m.weight.data)
    elif name.find('BatchNorm2d') != -1:
        init.normal_(m.weight.data, 1.0, 0.02)
        init.constant_(m.bias.data, 0.0)


def weights_init_normal(m):
    _init_by_name(m, lambda w: init.normal_(w, 0.0, 0.02))


def weights_init_xavier(m):
    _init_by_name(m, lambda w: init.xavier_normal_(w, gain=0.02))


def weights_init_kaiming(m):
    _init_by_name(m, lambda w: init.kaiming_normal_(w, a=0, mode='fan_in'))


def weights_init_orthogonal(m):
    _init_by_name(m, lambda w: init.orthogonal_(w, gain=1))


def init_weights(net, init_type='normal'):
    print('initialization method [%s]' % init_type)
    fns = {'normal': weights_init_normal, 'xavier': weights_init_xavier, 'kaiming': weights_init_kaiming,
           'orthogonal': weights_init_orthogonal}
    if init_type not in fns:
        raise NotImplementedError('initialization method [%s] is not implemented' % init_type)
    net.apply(fns[init_type])


def get_norm_layer(norm_type='instance'):
    if norm_type == 'batch':
        return functools.partial(nn.BatchNorm2d, affine=True)
    if norm_type == 'instance':
        return functools.partial(nn.InstanceNorm2d, affine=False)
    if norm_type == 'none':
        return None
    raise NotImplementedError('normalization layer [%s] is not found' % norm_type)


def get_scheduler(optimizer, opt):
    if opt.lr_policy == 'lambda':
        def lambda_rule(epoch):
            return 1.0 - max(0, epoch + 1 + opt.epoch_count - opt.niter) / float(opt.niter_decay + 1)
        return lr_scheduler.LambdaLR(optimizer, lr_lambda=lambda_rule)
    if opt.lr_policy == 'step':
        return lr_scheduler.StepLR(optimizer, step_size=opt.lr_decay_iters, gamma=0.1)
    if opt.lr_policy == 'plateau':
        return lr_scheduler.ReduceLROnPlateau(optimizer, mode='min', factor=0.2, threshold=0.01, patience=5)
    return NotImplementedError('learning rate policy [%s] is not implemented', opt.lr_policy)


def print_network(net):
    num_params = sum(p.numel() for p in net.parameters())
    print(net)
    print('Total number of parameters: %d' % num_params)


class _BceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, target):
        from mmhand_b200 import runtime
        ops = runtime.get_ops(x.device)
        xc = x.contiguous().float()
        acc = torch.zeros(1, dtype=torch.float32, device=x.device)
        grad = torch.empty_like(xc) if x.requires_grad else None
        n = xc.numel()
        ops.bce_logits(xc, target, 1.0 / n, 1.0 / n, acc, grad)
        ctx.grad = grad
        return acc[0]

    @staticmethod
    def backward(ctx, g):
        return (ctx.grad * g if ctx.grad is not None else None), None


class GANLoss(nn.Module):
    def __init__(self, use_lsgan=True, target_real_label=1.0, target_fake_label=0.0, gpu=0):
        super().__init__()
        self.real_label, self.fake_label = float(target_real_label), float(target_fake_label)
        self.device = gpu

    def get_target_tensor(self, prediction, target_is_real):
        v = self.real_label if target_is_real else self.fake_label
        return torch.tensor(v, device=prediction.device).expand_as(prediction)

    def __call__(self, input, target_is_real):
        return _BceFn.apply(input, self.real_label if target_is_real else self.fake_label)
