"""Host-side mirror of the reference's model surface for the ODE-Net hot path.

Same class names, constructor arguments, attribute names and therefore state_dict keys
as /root/reference/model.py (ODENet :6-62, downsamplers :119-228, FCClassifier :231-250,
ConcatConv2d :313-323, ODEfunc :326-348, ODEBlock :351-403), so checkpoints written by the
reference's train.py load unchanged and `torch.manual_seed(s); ODENet(...)` draws the same
random initial weights (modules are created in the same order).

The only compute that matters here is ODEBlock.forward -> odeint / odeint_adjoint, which is
served by the B200 kernels behind `torchdiffeq` (this repo's drop-in package).  Downsamplers
and the classifier run once per batch: their convolutions / pooling / linear stay plain PyTorch
(SURVEY 2, row 9); their GroupNorm -> ReLU pairs run as one CUDA pass when no gradient is needed
(node_b200/caller_ops.py).
"""
import torch
import torch.nn as nn

from .solver import odeint, odeint_adjoint
from .caller_ops import group_norm_relu, head, res_conv, res_head, run_sequential, stem_gn_relu


def _norm_factory(kind='group'):
    if kind == 'group':
        return lambda dim: nn.GroupNorm(min(32, dim), dim)
    if kind == 'batch':
        return lambda dim: nn.BatchNorm2d(dim, track_running_stats=False)
    raise NotImplementedError('Normalization layer not implemented: {}'.format(kind))


class ConcatConv2d(nn.Module):
    """Convolution over [t * ones, x]: the time plane is input channel 0."""

    def __init__(self, dim_in, dim_out, transpose=False, **kwargs):
        super().__init__()
        layer = nn.ConvTranspose2d if transpose else nn.Conv2d
        self._layer = layer(dim_in + 1, dim_out, **kwargs)

    def forward(self, t, x):
        plane = torch.ones_like(x[:, :1, :, :]) * t
        return self._layer(torch.cat([plane, x], 1))


class ODEfunc(nn.Module):
    """GN -> ReLU -> ConcatConv(t) -> GN -> ReLU -> ConcatConv(t) -> GN.

    forward() is the eager definition (used for autograd VJPs and as documentation); the
    solver recognises this structure and runs the fused sm_100a kernels instead, adding the
    same number of evaluations to `nfe`.
    """

    def __init__(self, dim, norm='group'):
        super().__init__()
        make = _norm_factory(norm)
        self.norm1 = make(dim)
        self.relu = nn.ReLU(inplace=True)
        self.conv1 = ConcatConv2d(dim, dim, kernel_size=3, stride=1, padding=1)
        self.norm2 = make(dim)
        self.conv2 = ConcatConv2d(dim, dim, kernel_size=3, stride=1, padding=1)
        self.norm3 = make(dim)
        self.nfe = 0

    def forward(self, t, x):
        self.nfe += 1
        h = self.relu(self.norm1(x))
        h = self.relu(self.norm2(self.conv1(t, h)))
        return self.norm3(self.conv2(t, h))


class ODEBlock(nn.Module):

    def __init__(self, n_filters=64, tol=1e-3, method='dopri5', adjoint=False, t1=1, norm='group'):
        super().__init__()
        self.odefunc = ODEfunc(n_filters, norm=norm)
        self.t1 = t1
        self.tol = tol
        self.method = method
        self.odeint = odeint_adjoint if adjoint else odeint
        self.return_last_only = True

    def forward(self, x):
        if self.integration_time is None:
            return x
        self.integration_time = self.integration_time.type_as(x)
        out = self.odeint(self.odefunc, x, self.integration_time, method=self.method,
                          rtol=self.tol, atol=self.tol)
        return out[-1] if self.return_last_only else out

    @property
    def nfe(self):
        return self.odefunc.nfe

    @nfe.setter
    def nfe(self, value):
        self.odefunc.nfe = value

    @property
    def t1(self):
        return self.integration_time[1]

    @t1.setter
    def t1(self, value):
        if isinstance(value, (int, float)):
            self.integration_time = None if value == 0 else torch.tensor([0, value], dtype=torch.float32)
            return
        if not isinstance(value, (list, tuple, torch.Tensor)):
            raise ValueError('Argument must be a scalar, a list, or a tensor')
        value = value.tolist() if isinstance(value, torch.Tensor) else list(value)
        if value[0] != 0:
            value = [0] + value
        self.integration_time = torch.tensor(value, dtype=torch.float32)


class Flatten(nn.Module):

    def forward(self, x):
        return x.reshape(x.shape[0], -1)


class ResBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, norm='group'):
        super().__init__()
        make = _norm_factory(norm)
        self.norm1 = make(inplanes)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.norm2 = make(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)

    def forward(self, x):
        out = group_norm_relu(self.norm1, x) if isinstance(self.norm1, nn.GroupNorm) else self.relu(self.norm1(x))
        return self.forward_act(out, x)

    def forward_act(self, out, x=None, next_norm=None):
        """The block after its first normalisation: out = relu(norm1(x)) (x is only needed for an identity shortcut).
        With `next_norm` (the following block's norm1) the result is relu(next_norm(block output))."""
        if self.downsample is not None and isinstance(self.norm1, nn.GroupNorm):
            out, shortcut = res_head(self.norm1, self.conv1, self.downsample, out)   # one kernel when served (caller_ops)
        else:
            shortcut = x if self.downsample is None else self.downsample(out)
            out = self.conv1(out)
        if isinstance(self.norm2, nn.GroupNorm):
            return res_conv(self.norm2, self.conv2, out, shortcut, next_norm)        # one kernel when served (caller_ops)
        out = self.conv2(self.relu(self.norm2(out))) + shortcut
        return out if next_norm is None else group_norm_relu(next_norm, out)


def _conv1x1(cin, cout, stride):
    return nn.Conv2d(cin, cout, kernel_size=1, stride=stride, bias=False)


class _Wrapped(nn.Module):
    def forward(self, *inputs):
        if isinstance(self.module, nn.Sequential) and len(inputs) == 1:
            return run_sequential(self.module, inputs[0])      # GroupNorm -> ReLU pairs in one pass (caller_ops)
        return self.module(*inputs)


class OneShotDownsample(_Wrapped):
    def __init__(self, in_ch, out_ch=64, **_):
        super().__init__()
        self.module = nn.Conv2d(in_ch, out_ch, 4, 2, 1)


class _ConvStack(_Wrapped):
    def __init__(self, in_ch, width, out_ch, norm):
        super().__init__()
        make = _norm_factory(norm)
        self.module = nn.Sequential(
            nn.Conv2d(in_ch, width, 3, 1), make(width), nn.ReLU(inplace=True),
            nn.Conv2d(width, width, 4, 2, 1), make(width), nn.ReLU(inplace=True),
            nn.Conv2d(width, out_ch, 4, 2, 1))


class MinimalConvDownsample(_ConvStack):
    def __init__(self, in_ch, out_ch=64, norm='group'):
        super().__init__(in_ch, 24, out_ch, norm)


class ConvDownsample(_ConvStack):
    def __init__(self, in_ch, out_ch=64, norm='group'):
        super().__init__(in_ch, 64, out_ch, norm)


class ResDownsample(_Wrapped):
    def __init__(self, in_ch, out_ch=64, norm='group'):
        super().__init__()
        self.module = nn.Sequential(
            nn.Conv2d(in_ch, 64, 3, 1),
            ResBlock(64, 64, stride=2, downsample=_conv1x1(64, 64, 2), norm=norm),
            ResBlock(64, out_ch, stride=2, downsample=_conv1x1(64, out_ch, 2), norm=norm))

    def forward(self, x):
        conv0, rb1, rb2 = self.module[0], self.module[1], self.module[2]
        if isinstance(rb1.norm1, nn.GroupNorm) and rb1.downsample is not None:
            # conv0's raw output feeds nothing but rb1.norm1 (the shortcut branches off AFTER the normalisation,
            # model.py:170-172): stem convolution, GroupNorm and ReLU in one pass (caller_ops.stem_gn_relu)
            a1 = stem_gn_relu(conv0, rb1.norm1, x)
            if isinstance(rb2.norm1, nn.GroupNorm) and rb2.downsample is not None:     # rb2 too reads only relu(norm1(.))
                return rb2.forward_act(rb1.forward_act(a1, next_norm=rb2.norm1))
            return rb2(rb1.forward_act(a1))
        return rb2(rb1(conv0(x)))


class ODEDownsample(nn.Module):
    def __init__(self, in_ch, out_ch=64, method='dopri5', adjoint=False, t1=1, tol=1e-3, norm='group'):
        super().__init__()
        self.conv1 = nn.Conv2d(in_ch, out_ch, 4, 2, 1)
        self.odeblock = ODEBlock(n_filters=out_ch, adjoint=adjoint, t1=t1, tol=tol, method=method, norm=norm)
        self.maxpool = nn.MaxPool2d(4, 2, 1)

    def forward(self, x):
        x = self.odeblock(self.conv1(x))
        if x.dim() > 4:
            return x, self.maxpool(x[-1])
        return self.maxpool(x)


class ODEDownsample2(nn.Module):
    def __init__(self, in_ch, out_ch=64, method='dopri5', adjoint=False, t1=1, tol=1e-3, norm='group'):
        super().__init__()
        self.conv1 = nn.Conv2d(in_ch, out_ch, 4, 2, 1)
        self.odeblock = ODEBlock(n_filters=out_ch, adjoint=adjoint, t1=t1, tol=tol, method=method, norm=norm)
        self.norm = nn.Sequential(_norm_factory(norm)(out_ch), nn.ReLU(inplace=True))
        self.conv2 = nn.Conv2d(out_ch, out_ch, 4, 2, 1)
        self.apply_conv = False

    def forward(self, x):
        x = self.odeblock(self.conv1(x))
        if x.dim() > 4:
            x = torch.stack([run_sequential(self.norm, xi) for xi in x])
            if self.apply_conv:
                x = torch.stack([self.conv2(xi) for xi in x])
                return x, x[-1]
            return x, self.conv2(x[-1])
        return self.conv2(run_sequential(self.norm, x))


class FCClassifier(_Wrapped):
    def __init__(self, in_ch=64, out=10, dropout=0, norm='group'):
        super().__init__()
        layers = [_norm_factory(norm)(in_ch), nn.ReLU(inplace=True), nn.AdaptiveAvgPool2d((1, 1))]
        if dropout:
            layers.append(nn.Dropout(dropout))
        layers += [Flatten(), nn.Linear(in_ch, out)]
        self.module = nn.Sequential(*layers)

    def forward(self, x):
        return head(self.module, x)          # GroupNorm -> ReLU -> pool -> linear in one pass when served (caller_ops)


_DOWNSAMPLERS = {
    'residual': ResDownsample, 'convolution': ConvDownsample,
    'minimal': MinimalConvDownsample, 'one-shot': OneShotDownsample,
}


class ODENet(nn.Module):

    def __init__(self, in_ch, out=10, n_filters=64, downsample='residual', method='dopri5', tol=1e-3,
                 adjoint=False, t1=1, dropout=0, norm='group'):
        super().__init__()
        if downsample in _DOWNSAMPLERS:
            self.downsample = _DOWNSAMPLERS[downsample](in_ch, out_ch=n_filters, norm=norm)
        elif downsample in ('ode', 'ode2'):
            cls = ODEDownsample if downsample == 'ode' else ODEDownsample2
            self.downsample = cls(in_ch, out_ch=n_filters, norm=norm, adjoint=adjoint, t1=t1, tol=tol, method=method)
        self.odeblock = ODEBlock(n_filters=n_filters, tol=tol, adjoint=adjoint, t1=t1, method=method, norm=norm)
        self.classifier = FCClassifier(in_ch=n_filters, out=out, dropout=dropout, norm=norm)

    def forward(self, x):
        out = []
        x = self.downsample(x)
        if isinstance(x, (tuple, list)):
            feats, x = x
            if isinstance(self.classifier.module[-1], nn.Sequential):
                feats = torch.stack([f.mean(-1).mean(-1) for f in feats])
            else:
                feats = torch.stack([self.classifier(f) for f in feats])
            out.append(feats)
        x = self.odeblock(x)
        if x.dim() > 4:
            x = torch.stack([self.classifier(xi) for xi in x])
        else:
            x = self.classifier(x)
        out.append(x)
        return torch.cat(out)

    def to_features_extractor(self, keep_pool=True):
        if isinstance(self.downsample, (ODEDownsample, ODEDownsample2)):
            self.downsample.odeblock.return_last_only = False
        self.odeblock.return_last_only = False
        if keep_pool:
            self.classifier.module[-1] = nn.Sequential()
        else:
            self.classifier = nn.Sequential(*list(self.classifier.module.children())[:2])

    def nfe(self, reset=False):
        n = self.odeblock.nfe
        if reset:
            self.odeblock.nfe = 0
        return n
