"""Parameter-holding leaf modules. They reproduce the reference's ``state_dict`` names, shapes and dtypes
(fp32, PyTorch layouts) -- models/Generator.py and models/Discriminator.py build nn.Sequential containers whose
slot indices are part of the checkpoint ABI -- but hold no compute: the arithmetic runs
in mmhand_b200.engine on the CUDA kernels. Class names contain 'Conv' / 'BatchNorm2d' so that the reference's
``init_weights`` (models/network_utils.py:12-21, matching on class names) initialises them identically.
"""
import torch
import torch.nn as nn


class Conv2dParams(nn.Module):
    def __init__(self, cin, cout, k, bias=False, stride=1):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size, self.stride = cin, cout, k, stride
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k))
        self.bias = nn.Parameter(torch.empty(cout)) if bias else None
        # nn.Conv2d default init (kaiming_uniform(a=sqrt(5))); overwritten by init_weights in the training path
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        if bias:
            bound = 1.0 / (cin * k * k) ** 0.5
            nn.init.uniform_(self.bias, -bound, bound)

    def extra_repr(self):
        return "%d, %d, kernel_size=%d, stride=%d, bias=%s" % (self.in_channels, self.out_channels, self.kernel_size,
                                                               self.stride, self.bias is not None)


class ConvTranspose2dParams(nn.Module):
    def __init__(self, cin, cout, k, bias=False):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = cin, cout, k
        self.weight = nn.Parameter(torch.empty(cin, cout, k, k))
        self.bias = nn.Parameter(torch.empty(cout)) if bias else None
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        if bias:
            nn.init.zeros_(self.bias)

    def extra_repr(self):
        return "%d, %d, kernel_size=%d, stride=2" % (self.in_channels, self.out_channels, self.kernel_size)


class BatchNorm2dParams(nn.Module):
    def __init__(self, c, affine=True):
        super().__init__()
        assert affine, "BatchNorm2d(affine=True) is the only normalisation the reference's options produce"
        self.num_features = c
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        self._pending = 0

    def note_batch(self):
        self._pending += 1          # host-side counter, folded into the buffer when the state is read

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        if self._pending:
            self.num_batches_tracked += self._pending
            self._pending = 0
        super()._save_to_state_dict(destination, prefix, keep_vars)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        self._pending = 0           # batches counted before the load belong to the overwritten value
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def extra_repr(self):
        return "%d" % self.num_features


class Slot(nn.Module):
    """Placeholder that keeps the reference's nn.Sequential numbering (padding / ReLU / Dropout / Tanh slots)."""

    def __init__(self, what):
        super().__init__()
        self.what = what

    def extra_repr(self):
        return self.what


def norm_params(norm_layer, c):
    """The module that sits in the norm slot: BatchNorm2d parameters, or -- for the reference's
    ``partial(InstanceNorm2d, affine=False)`` (network_utils.py get_norm_layer) -- a parameter-free placeholder: such a
    model has the reference's state_dict (biased convolutions, no norm entries) and loads its checkpoints, but the
    B200 path computes batch normalisation only and refuses to run it (forward raises)."""
    if norm_kind(norm_layer) == 'instance':
        return Slot('InstanceNorm2d(%d, affine=False)' % c)
    return BatchNorm2dParams(c)


def norm_kind(norm_layer):
    """'batch' | 'instance' from the reference's norm_layer argument (a class or a functools.partial)."""
    import functools
    f = norm_layer.func if isinstance(norm_layer, functools.partial) else norm_layer
    if f is nn.BatchNorm2d:
        return 'batch'
    if f is nn.InstanceNorm2d:
        return 'instance'
    raise NotImplementedError("normalization layer %r is not supported" % (norm_layer,))
