"""RC-Net networks, host side (mirror of the reference's ``RCNet/networks.py``).

The CNN backbones stay plain torch modules (BASELINE.json north_star: "the CNN
backbones ... remain the reference torch modules"); they are re-declared here,
table-driven, with the reference's attribute names so that reference
checkpoints (``radarnet_encoder_state_dict`` / ``radarnet_decoder_state_dict``,
RCNet/rcnet_model.py:211-257) load unchanged:

  ResNetEncoder        <- RCNet/networks.py:10-270   (ResNet-18/34 trunk, skips at 1/2..1/16)
  FullyConnectedEncoder<- RCNet/networks.py:273-332  (6 FC layers, LeakyReLU 0.2)
  RCNetEncoder         <- RCNet/networks.py:335-451  (backbone + gather + point MLP + LoFTR)
  MultiScaleDecoder    <- RCNet/networks.py:458-778  (nearest-up + conv blocks with skips)
  building blocks      <- utils/net_utils.py:29-569

What is NOT torch here is the hot path inside ``RCNetEncoder.forward``: the five
``torchvision.ops.roi_pool`` calls (networks.py:418-433) become one launch of the
sm_100a gather kernel, and the LoFTR transformer (networks.py:444) runs through
the fused layer kernels.  There is no CPU fallback for those two.
"""
import functools
import os

import torch
import torch.nn.functional as F

from .linear_attention import LocalFeatureTransformer
from .. import ops

_LEAK = 0.20   # utils/net_utils.py:15  LeakyReLU(negative_slope=0.20)
# form of the sub-pixel phase convolution of the exact-2x / 2n+1 decoder blocks (DecoderBlock.fold_upsample_): 3 = 3x3 phase
# kernels with their zero taps, 2 = 2x2 phase kernels (same sums, 4/9 of the multiply-adds)
SUBPIXEL_TAPS = int(os.environ.get('RIDERS_SUBPIXEL_TAPS', '2'))


def _init_weight(weight, how):
    # utils/net_utils.py:61-66: torch default (kaiming_uniform a=sqrt(5)) unless named otherwise
    if how == 'kaiming_normal':
        torch.nn.init.kaiming_normal_(weight)
    elif how == 'xavier_normal':
        torch.nn.init.xavier_normal_(weight)
    elif how == 'xavier_uniform':
        torch.nn.init.xavier_uniform_(weight)


def _act(name):
    """utils/net_utils.py:4-26 activation lookup; returns a callable or None."""
    if name is None or 'linear' in name:
        return None
    if 'leaky_relu' in name:
        return functools.partial(F.leaky_relu, negative_slope=_LEAK, inplace=True)    # picklable / deep-copyable
    if 'relu' in name:
        return torch.relu
    if 'elu' in name:
        return F.elu
    if 'sigmoid' in name:
        return torch.sigmoid
    raise ValueError('Unsupported activation function: {}'.format(name))


class Conv2d(torch.nn.Module):
    """conv (no bias) -> optional BatchNorm -> activation; utils/net_utils.py:29-81."""

    def __init__(self, cin, cout, kernel_size=3, stride=1, init='kaiming_uniform',
                 act='leaky_relu', use_batch_norm=False):
        super().__init__()
        self.conv = torch.nn.Conv2d(cin, cout, kernel_size, stride, kernel_size // 2, bias=False)
        _init_weight(self.conv.weight, init)
        self.use_batch_norm = use_batch_norm
        if use_batch_norm:
            self.batch_norm = torch.nn.BatchNorm2d(cout)
        self._act = _act(act)
        self._act_name = None if (act is None or 'linear' in act) else ('leaky_relu' if 'leaky_relu' in act else
                                                                        ('relu' if 'relu' in act else act))
        self._fused_slope = None                  # set by fold_batch_norm_(): bias + activation run in one kernel

    def forward(self, x):
        if self._fused_slope is not None and x.is_cuda and not torch.is_grad_enabled() and x.dtype == torch.float32 \
                and (ops._is_channels_last(x) or x.shape[1] == 1):
            # inference fast path after fold_batch_norm_(): conv without bias, then bias + activation in ONE pass
            c = self.conv
            y = F.conv2d(x, c.weight, None, c.stride, c.padding, c.dilation, c.groups)
            if c.bias is None and self._fused_slope == 1.0:
                return y
            if ops._is_channels_last(y) or (y.shape[1] == 1 and y.is_contiguous()):
                return ops.bias_act_(y, c.bias, self._fused_slope)
            y = y if c.bias is None else y + c.bias.view(1, -1, 1, 1)
            return y if self._act is None else self._act(y)
        y = self.conv(x)
        if self.use_batch_norm:
            y = self.batch_norm(y)
        return y if self._act is None else self._act(y)

    def forward_padded(self, x, sy, sx):
        """Inference fast path only (needs ``fast_ok(x)``): act(conv(x) + bias) replicate-padded by one row on top (sy) /
        one column on the left (sx), written by the bias + activation pass itself.  Returns the padded map; the
        activated map is ``out[:, :, sy:, sx:]``."""
        c = self.conv
        y = F.conv2d(x, c.weight, None, c.stride, c.padding, c.dilation, c.groups)
        return ops.bias_act_pad(y, c.bias, self._fused_slope, sy, sx)

    def fast_ok(self, x):
        return (self._fused_slope is not None and x.is_cuda and not torch.is_grad_enabled() and x.dtype == torch.float32
                and ops._is_channels_last(x))

    def raw(self, x):
        """Convolution only; the caller fuses (bias, slope) into its own kernel.  Needs ``fast_ok(x)``."""
        c = self.conv
        return F.conv2d(x, c.weight, None, c.stride, c.padding, c.dilation, c.groups), c.bias, self._fused_slope

    def fold_batch_norm_(self):
        """Inference only: fold the (eval-mode) BatchNorm affine into the convolution's weights and bias
        (torch.nn.utils.fusion).  Same function up to float32 rounding; removes one full pass over the
        activations per convolution.  The state_dict no longer matches the reference's after this."""
        if self.use_batch_norm:
            self.conv = torch.nn.utils.fusion.fuse_conv_bn_eval(self.conv.eval(), self.batch_norm.eval())
            self.use_batch_norm = False
            del self.batch_norm
        self._fused_slope = {'leaky_relu': _LEAK, 'relu': 0.0, None: 1.0}.get(self._act_name)
        return self


class UpConv2d(torch.nn.Module):
    """nearest resize to ``shape`` then Conv2d; utils/net_utils.py:149-190."""

    def __init__(self, cin, cout, init, act, use_batch_norm):
        super().__init__()
        self.conv = Conv2d(cin, cout, 3, 1, init, act, use_batch_norm)

    def forward(self, x, shape):
        if x.is_cuda and x.dtype == torch.float32 and x.shape[1] % 4 == 0 and ops._is_channels_last(x) \
                and not torch.is_grad_enabled():
            return self.conv(ops.upsample_nearest(x, shape))      # same values, streaming NHWC kernel
        return self.conv(F.interpolate(x, size=shape))


class FullyConnected(torch.nn.Module):
    """Linear (+bias) -> activation; utils/net_utils.py:193-236."""

    def __init__(self, fin, fout, init='kaiming_uniform', act='leaky_relu'):
        super().__init__()
        self.fully_connected = torch.nn.Linear(fin, fout)
        _init_weight(self.fully_connected.weight, init)
        self._act = _act(act)

    def forward(self, x):
        y = self.fully_connected(x)
        return y if self._act is None else self._act(y)


class ResNetBlock(torch.nn.Module):
    """Two 3x3 convs + (1x1 projection when shape/channels change); utils/net_utils.py:242-310."""

    def __init__(self, cin, cout, stride, init, act, use_batch_norm):
        super().__init__()
        self.conv1 = Conv2d(cin, cout, 3, stride, init, act, use_batch_norm)
        self.conv2 = Conv2d(cout, cout, 3, 1, init, act, use_batch_norm)
        # the reference always owns a projection (it is in every checkpoint) even when unused
        self.projection = Conv2d(cin, cout, 1, stride, init, None, False)
        self._needs_projection = stride != 1 or cin != cout
        self._act = _act(act)

    def forward(self, x):
        y = self.conv2(self.conv1(x))
        skip = self.projection(x) if self._needs_projection else x
        return self._act(y + skip)


class DecoderBlock(torch.nn.Module):
    """up-conv, concat skip, conv; utils/net_utils.py:474-569 (deconv_type 'up' only)."""

    def __init__(self, cin, cskip, cout, init, act, use_batch_norm, deconv_type='up'):
        super().__init__()
        if deconv_type != 'up':
            raise ValueError("only deconv_type='up' is used by RC-Net (rcnet_model.py:94)")
        self.skip_channels = cskip
        # sub-pixel weights, built by fold_upsample_(): non-persistent buffers, so .to(device) / .half() move them with the
        # module while the state_dict keeps the reference's keys only
        for name in ('_w4', '_w2', '_wcol', '_wrow'):
            self.register_buffer(name, None, persistent=False)
        self.subpixel_taps = SUBPIXEL_TAPS          # 3: 3x3 phase kernels; 2: the equivalent 2x2 phase kernels (4/9 of the MACs)
        self._leaky = act is not None and 'leaky_relu' in act
        # set by MultiScaleDecoder.forward: may this block's up-convolution run as ONE tcgen05 kernel (ops.upconv_tc_f32, 64 -> 32
        # channels only), and is its fp16-range flag collected for the caller's sync point instead of being read here
        self.upconv_tc, self.upconv_defer, self.last_used_upconv_tc = False, False, False
        self.deconv = UpConv2d(cin, cout, init, act, use_batch_norm)
        self.conv = Conv2d(cskip + cout, cout, 3, 1, init, act, use_batch_norm)

    def fold_upsample_(self):
        """Inference only, after the BatchNorm fold: weights of the sub-pixel form of 'nearest 2x resize -> 3x3 conv'.
        Output pixel (2i+a, 2j+b) of conv3x3(resize2x(x)) reads x rows {i-1, i, i+1} with taps [W0, W1+W2, 0] (a = 0) or
        [0, W0+W1, W2] (a = 1), likewise for columns; the four phases become 4*Cout output channels, phase-major."""
        c = self.deconv.conv
        if c.use_batch_norm or not self._leaky or c.conv.kernel_size != (3, 3) or c.conv.out_channels % 4:
            return self
        w = c.conv.weight.detach()
        t = [torch.tensor([[1., 0., 0.], [0., 1., 1.], [0., 0., 0.]], device=w.device, dtype=w.dtype),
             torch.tensor([[0., 0., 0.], [1., 1., 0.], [0., 0., 1.]], device=w.device, dtype=w.dtype)]
        w4 = torch.cat([torch.einsum('sk,oikl,tl->oist', t[a], w, t[b]) for a in (0, 1) for b in (0, 1)], dim=0)
        self._w4 = w4.contiguous(memory_format=torch.channels_last)
        # the zero taps dropped: phase a reads rows {i-1+a, i+a} with [W0, W1+W2] (a = 0) / [W0+W1, W2] (a = 1); as ONE 2x2
        # convolution with padding 1 the phase (a, b) of pixel (i, j) is output pixel (i+a, j+b) of an (h+1, w+1) map
        t2 = [torch.tensor([[1., 0., 0.], [0., 1., 1.]], device=w.device, dtype=w.dtype),
              torch.tensor([[1., 1., 0.], [0., 0., 1.]], device=w.device, dtype=w.dtype)]
        w2 = torch.cat([torch.einsum('sk,oikl,tl->oist', t2[a], w, t2[b]) for a in (0, 1) for b in (0, 1)], dim=0)
        self._w2 = w2.contiguous(memory_format=torch.channels_last)
        # 2n+1 targets (torch nearest picks source max(j-1, 0) >> 1, see upsample.cu): the first output column (row) sees the
        # zero padding on its left (top), so its kx (ky) in {1, 2} taps are pre-summed into a 3x1 (1x3) phase kernel
        wc, wr = w[:, :, :, 1] + w[:, :, :, 2], w[:, :, 1, :] + w[:, :, 2, :]
        self._wcol = torch.cat([torch.einsum('sk,oik->ois', t[a], wc) for a in (0, 1)], dim=0).unsqueeze(-1).contiguous()
        self._wrow = torch.cat([torch.einsum('tl,oil->oit', t[b], wr) for b in (0, 1)], dim=0).unsqueeze(2).contiguous()
        return self

    def _tail(self, z, raw, pad_out):
        if raw:
            return self.conv.raw(z)
        if pad_out is not None and self.conv.fast_ok(z) and self.conv.conv.out_channels % 4 == 0:
            yp = self.conv.forward_padded(z, pad_out[0], pad_out[1])
            return yp[:, :, pad_out[0]:, pad_out[1]:], yp
        return self.conv(z)

    def upconv_tc_ok(self, x_shape, target):
        """Will ``forward`` run this block's up-convolution as ONE tcgen05 kernel (ops.upconv_tc_f32) for an input of ``x_shape``
        and the target size ``target``?  (64 -> 32 channels, exact 2x, inference, TF32 flow; decided by MultiScaleDecoder.)"""
        return (self.upconv_tc and self.subpixel_taps == 2 and self._w2 is not None and tuple(self._w2.shape) == (128, 64, 2, 2)
                and x_shape[1] == 64 and x_shape[3] <= 26 and tuple(target) == (2 * x_shape[2], 2 * x_shape[3])
                and torch.backends.cudnn.allow_tf32 and not torch.is_grad_enabled())

    def forward(self, x, skip=None, shape=None, raw=False, xp=None, pad_out=None, x_act=None):
        """``raw=True`` (inference fast path only): return (conv output without bias/activation, bias, slope)
        of the block's last convolution so that the decoder head can fuse them.
        ``x_act=(bias, slope)``: ``x`` is the RAW output of the previous block's convolution (its bias + LeakyReLU are applied by
        ops.upconv_tc_f32 while it stages its operand, or here in one pass when that kernel does not take this call).
        ``pad_out=(sy, sx)`` (fast path only, ignored otherwise): the NEXT block's target is 2n+1 along one axis, so return
        ``(y, y_padded)`` with ``y = y_padded[:, :, sy:, sx:]``; ``xp``: that padded map handed to the next block."""
        if skip is not None:
            shape = skip.shape[2:4]
        elif shape is None:
            shape = (2 * x.shape[2], 2 * x.shape[3])
        up = self.deconv.conv
        taps = 2 if (self.subpixel_taps == 2 and self._w2 is not None) else 3
        wsub = self._w2 if taps == 2 else self._w4
        use_upconv = (self.upconv_tc_ok(x.shape, shape) and ops._is_channels_last(x) and x.is_cuda and x.dtype == torch.float32
                      and (skip is None or skip.shape[1] % 4 == 0))

        def activate(t):
            if x_act is None or (x_act[0] is None and float(x_act[1]) == 1.0):
                return t
            if ops._is_channels_last(t) and t.is_cuda:
                return ops.bias_act_(t, x_act[0], x_act[1])
            return F.leaky_relu(t if x_act[0] is None else t + x_act[0].view(1, -1, 1, 1), float(x_act[1]))
        if x_act is not None and not use_upconv:
            x, x_act = activate(x), None
        # a skip gathered straight into this block's concat buffer (ops.placed_skip_buffers) counts as channels-last
        skip_cl = skip is not None and (ops._is_channels_last(skip) or ops.placed_parent(skip, up.conv.out_channels) is not None)
        if (self._w4 is not None and tuple(shape) == (2 * x.shape[2], 2 * x.shape[3]) and up.fast_ok(x)
                and (skip is None or (skip_cl and skip.shape[1] % 4 == 0))):
            # exact 2x block: sub-pixel convolution on the low-resolution map (the resized map never exists),
            # then depth-to-space + bias + LeakyReLU (+ skip concat) in one pass
            if use_upconv:
                # TF32 flow, 64 -> 32 channels (RC-Net's deconv1): (previous bias + LeakyReLU +) phase convolution + depth-to-space
                # + bias + LeakyReLU + concat in ONE tcgen05 kernel (csrc/upconv_tc.cu); fp16 operands, the range flag joins
                # the caller's pending guards
                xb, xs = (None, 1.0) if x_act is None else x_act
                z, flag = ops.upconv_tc_f32(x, xb, xs, wsub, up.conv.bias, _LEAK, skip if self.skip_channels > 0 else None)
                self.last_used_upconv_tc = True
                if self.upconv_defer:
                    ops.LoftrOverflow.pending.append((ops.TailOverflow(flag, None), (z,)))
                elif int(flag.item()):                                 # a value left the fp16 range: the library passes answer
                    self.last_used_upconv_tc = False
                    y4 = F.conv2d(activate(x), wsub, None, 1, 1)
                    z = ops.d2s_bias_act_concat(y4, up.conv.bias, skip if self.skip_channels > 0 else None, _LEAK, taps=taps)
                return self._tail(z, raw, pad_out)
            y4 = F.conv2d(x, wsub, None, 1, 1)
            z = ops.d2s_bias_act_concat(y4, up.conv.bias, skip if self.skip_channels > 0 else None, _LEAK, taps=taps)
            return self._tail(z, raw, pad_out)
        sy, sx = int(shape[0] == 2 * x.shape[2] + 1), int(shape[1] == 2 * x.shape[3] + 1)
        if (self._w4 is not None and sy + sx == 1 and shape[0] == 2 * x.shape[2] + sy and shape[1] == 2 * x.shape[3] + sx
                and up.fast_ok(x if xp is None else xp) and x.shape[2] > 1 and x.shape[3] > 1
                and (skip is None or (skip_cl and skip.shape[1] % 4 == 0))):
            # 2n+1 along one axis: the same sub-pixel convolution on x replicate-padded by one, read one pixel further, plus a
            # 3x1 / 1x3 phase convolution of the first column / row (the resized map never exists here either)
            if xp is None or tuple(xp.shape[2:]) != (x.shape[2] + sy, x.shape[3] + sx):
                if ops._is_channels_last(x) and x.shape[1] % 4 == 0:
                    xp = ops.bias_act_pad(x, None, 1.0, sy, sx)        # identity pass = replicate pad, one streaming kernel
                else:
                    xp = F.pad(x, (sx, 0, sy, 0), mode='replicate').contiguous(memory_format=torch.channels_last)
            y4 = F.conv2d(xp, wsub, None, 1, 1)
            if sx:
                yfirst = F.conv2d(x[:, :, :, 0:1].contiguous(memory_format=torch.channels_last), self._wcol, None, 1, (1, 0))
            else:
                yfirst = F.conv2d(x[:, :, 0:1, :].contiguous(memory_format=torch.channels_last), self._wrow, None, 1, (0, 1))
            yfirst = yfirst.contiguous(memory_format=torch.channels_last)
            z = ops.d2s_shift_bias_act_concat(y4, yfirst, up.conv.bias, skip if self.skip_channels > 0 else None, sy, sx, _LEAK,
                                              taps=taps)
            return self._tail(z, raw, pad_out)
        if xp is not None:
            x = x.contiguous(memory_format=torch.channels_last)
        if raw:
            return self.conv.raw(self.deconv(x, shape=shape))
        if (self.skip_channels > 0 and self._leaky and not up.use_batch_norm and not torch.is_grad_enabled()
                and x.is_cuda and x.dtype == torch.float32 and ops._is_channels_last(x) and skip_cl
                and x.shape[1] % 4 == 0 and up.conv.out_channels % 4 == 0 and skip.shape[1] % 4 == 0):
            # inference fast path (BatchNorm folded): resize -> conv(+bias) -> [LeakyReLU | skip] in one pass -> conv
            c = up.conv
            y = F.conv2d(ops.upsample_nearest(x, shape), c.weight, None, c.stride, c.padding, c.dilation, c.groups)
            return self.conv(ops.leaky_concat(y, skip, _LEAK, bias=c.bias))
        y = self.deconv(x, shape=shape)
        if self.skip_channels > 0:
            y = torch.cat([y, skip], dim=1)
        return self.conv(y)


class ResNetEncoder(torch.nn.Module):
    """7x7/2 stem, 3x3/2 max-pool, then ``blocks2..`` stages; RCNet/networks.py:10-270."""

    def __init__(self, n_layer, input_channels=3, n_filters=(32, 64, 128, 256, 256),
                 weight_initializer='kaiming_uniform', activation_func='leaky_relu',
                 use_batch_norm=False):
        super().__init__()
        depth = {18: [2, 2, 2, 2], 34: [3, 4, 6, 3]}.get(n_layer)
        if depth is None:
            raise ValueError('Only supports 18, 34 layer architecture')
        n_filters = list(n_filters)
        depth = depth + [depth[-1]] * (len(n_filters) - len(depth) - 1)
        assert len(n_filters) < 8 and len(n_filters) == len(depth) + 1
        a, w, bn = activation_func, weight_initializer, use_batch_norm
        self.conv1 = Conv2d(input_channels, n_filters[0], 7, 2, w, a, bn)
        self.max_pool = torch.nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self._stages = []
        for s in range(6):                      # blocks2 .. blocks7
            name = 'blocks%d' % (s + 2)
            if s + 1 < len(n_filters):
                cin, cout = n_filters[s], n_filters[s + 1]
                stride = 1 if s == 0 else 2
                blocks = [ResNetBlock(cin if n == 0 else cout, cout, stride if n == 0 else 1, w, a, bn)
                          for n in range(depth[s])]
                setattr(self, name, torch.nn.Sequential(*blocks))
                self._stages.append(name)
            else:
                setattr(self, name, None)

    def forward(self, x):
        pad = getattr(self, 'stem_pad_channels', 0)            # set by RCNetModel.optimize_for_inference()
        if pad and x.shape[1] + pad == self.conv1.conv.in_channels:
            xp = torch.empty((x.shape[0], x.shape[1] + pad) + tuple(x.shape[2:]), device=x.device, dtype=x.dtype,
                             memory_format=torch.channels_last)
            xp[:, :x.shape[1]].copy_(x)
            xp[:, x.shape[1]:].zero_()
            x = xp
        feats = [self.conv1(x)]
        y = self.max_pool(feats[0])
        for name in self._stages:
            y = getattr(self, name)(y)
            feats.append(y)
        return feats[-1], feats[:-1]


class FullyConnectedEncoder(torch.nn.Module):
    """RCNet/networks.py:273-332: in -> n_neurons[0..4] -> latent_size, activation after every layer."""

    def __init__(self, input_channels=3, n_neurons=(32, 64, 96, 128, 256), latent_size=29 * 10,
                 weight_initializer='kaiming_uniform', activation_func='leaky_relu'):
        super().__init__()
        widths = [input_channels] + list(n_neurons[:5]) + [latent_size]
        self.mlp = torch.nn.Sequential(*[
            FullyConnected(widths[i], widths[i + 1], weight_initializer, activation_func)
            for i in range(6)])

    def forward(self, x):
        return self.mlp(x)


class RCNetEncoder(torch.nn.Module):
    """Radar association encoder; RCNet/networks.py:335-451."""

    def __init__(self, input_channels_image=3, input_channels_depth=3,
                 input_patch_size_image=(900, 288),
                 n_filters_encoder_image=(32, 64, 128, 128, 128),
                 n_neurons_encoder_depth=(32, 64, 128, 128, 128),
                 latent_size_depth=128 * 29 * 10,
                 weight_initializer='kaiming_uniform', activation_func='leaky_relu',
                 use_batch_norm=False):
        super().__init__()
        self.n_neuron_latent_depth = n_neurons_encoder_depth[-1]
        self.encoder_image = ResNetEncoder(18, input_channels_image, n_filters_encoder_image,
                                           weight_initializer, activation_func, use_batch_norm)
        self.attention = LocalFeatureTransformer(['self', 'cross'], n_layers=4,
                                                 d_model=self.n_neuron_latent_depth)
        self.encoder_depth = FullyConnectedEncoder(input_channels_depth, n_neurons_encoder_depth,
                                                   latent_size_depth, weight_initializer, activation_func)
        self.input_patch_size_image = input_patch_size_image

    def pyramid(self, n_skips):
        """(scale, (oh, ow)) per gathered level, skips first then latent; networks.py:399-412."""
        ph, pw = self.input_patch_size_image[-2], self.input_patch_size_image[-1]
        scales = [1 / 2.0, 1 / 4.0, 1 / 8.0, 1 / 16.0, 1 / 32.0, 1 / 64.0, 1 / 128.0]
        levels = [(s, (int(ph * s), int(pw * s))) for s in scales[:n_skips]]
        levels.append((1 / 32.0, (int(ph // 32.0), int(pw // 32.0))))
        return levels

    def forward(self, image, points, b_boxes):
        """image (B,3,H',W'); points (sum K, 3); b_boxes list[B] of (K,4) -> latent, pooled skips."""
        latent_image, skips_image = self.encoder_image(image)
        return self.encode_points(latent_image, skips_image, points, b_boxes)

    def encode_latent(self, latent_image, points, b_boxes):
        """Latent half of ``encode_points``: gather of the 1/32 level, point MLP, LoFTR fusion, concat
        (networks.py:418-419, 438-450).  Cheap per point, so a whole batch goes through in one call."""
        scale, osize = self.pyramid(0)[-1]
        ph, pw = self.input_patch_size_image[-2], self.input_patch_size_image[-1]
        latent_pooled = ops.roi_gather([latent_image], b_boxes, [scale], [osize], max_roi=(pw, ph))[0]
        n, c = points.shape[0], self.n_neuron_latent_depth
        latent_depth = self.encoder_depth(points).view(n, c, -1, osize[1])
        tok_depth = latent_depth.reshape(n, c, -1).permute(0, 2, 1)
        tok_image = latent_pooled.reshape(n, latent_pooled.shape[1], -1).permute(0, 2, 1)
        tok_depth, tok_image = self.attention(tok_depth, tok_image)
        depth_tf = tok_depth.permute(0, 2, 1).reshape(latent_depth.shape)
        image_tf = tok_image.permute(0, 2, 1).reshape(latent_pooled.shape)
        return torch.cat([image_tf, depth_tf], dim=1)

    def gather_skips(self, skips_image, b_boxes, lead_channels=None):
        """Skip half of ``encode_points``: the four pooled skip levels (networks.py:421-433); 1.38 MB per
        point, so callers chunk the points to bound the decoder's activation memory.  ``lead_channels`` (one per
        level, from ``MultiScaleDecoder.skip_lead_channels``): gather each level straight into the concat buffer of
        the decoder block that consumes it, behind that many free channels (channels-last maps only)."""
        levels = self.pyramid(len(skips_image))[:-1]
        ph, pw = self.input_patch_size_image[-2], self.input_patch_size_image[-1]
        outs = None
        if (lead_channels is not None and all(ops._is_channels_last(f) for f in skips_image)
                and all(l > 0 and l % 4 == 0 for l in lead_channels) and torch.is_tensor(b_boxes)):
            outs = ops.placed_skip_buffers(b_boxes.shape[0], lead_channels, [f.shape[1] for f in skips_image],
                                           [o for _, o in levels], skips_image[0].device)
        return ops.roi_gather(list(skips_image), b_boxes, [s for s, _ in levels], [o for _, o in levels], max_roi=(pw, ph),
                              outs=outs)

    def encode_points(self, latent_image, skips_image, points, b_boxes):
        """Everything after the backbone (networks.py:418-451).  ``b_boxes`` may also be an (R,5)
        RoI tensor (batch index first), which lets a caller run chunks of points of a big batch."""
        levels = self.pyramid(len(skips_image))
        ph, pw = self.input_patch_size_image[-2], self.input_patch_size_image[-1]
        # one call for all levels (replaces the five roi_pool calls, networks.py:418-433)
        feats = list(skips_image) + [latent_image]
        if torch.is_grad_enabled() and any(f.requires_grad for f in feats):
            pooled = ops.roi_gather_autograd(feats, b_boxes, [s for s, _ in levels], [o for _, o in levels])   # training
        else:
            pooled = ops.roi_gather(feats, b_boxes, [s for s, _ in levels], [o for _, o in levels], max_roi=(pw, ph))
        skips_pooled, latent_pooled = pooled[:-1], pooled[-1]

        n, c = points.shape[0], self.n_neuron_latent_depth
        lw = levels[-1][1][1]
        latent_depth = self.encoder_depth(points).view(n, c, -1, lw)
        tok_depth = latent_depth.reshape(n, c, -1).permute(0, 2, 1)
        tok_image = latent_pooled.reshape(n, latent_pooled.shape[1], -1).permute(0, 2, 1)
        tok_depth, tok_image = self.attention(tok_depth, tok_image)
        depth_tf = tok_depth.permute(0, 2, 1).reshape(latent_depth.shape)
        image_tf = tok_image.permute(0, 2, 1).reshape(latent_pooled.shape)
        return torch.cat([image_tf, depth_tf], dim=1), skips_pooled


class MultiScaleDecoder(torch.nn.Module):
    """RCNet/networks.py:458-778.  RC-Net always builds it with n_resolution=1 and
    output_func='linear' (rcnet_model.py:83-95); that configuration (and deeper
    n_filters lists) is what is mirrored - the multi-resolution side outputs are not."""

    def __init__(self, input_channels=256, output_channels=1, n_resolution=1,
                 n_filters=(256, 128, 64, 32, 16), n_skips=(256, 128, 64, 32, 0),
                 weight_initializer='kaiming_uniform', activation_func='leaky_relu',
                 output_func='linear', use_batch_norm=False, deconv_type='up'):
        super().__init__()
        if n_resolution != 1 or 'upsample' in output_func:
            raise ValueError('only the single-resolution decoder used by RC-Net is provided')
        depth = len(n_filters)
        assert depth < 8, 'Does not support network depth of 8 or more'
        self.n_resolution = n_resolution
        cin = input_channels
        self._order = []
        for i in range(depth):                   # deconv{depth-1} ... deconv0
            name = 'deconv%d' % (depth - 1 - i)
            setattr(self, name, DecoderBlock(cin, n_skips[i], n_filters[i], weight_initializer,
                                             activation_func, use_batch_norm, deconv_type))
            self._order.append(name)
            cin = n_filters[i]
        for i in range(depth, 7):
            setattr(self, 'deconv%d' % i, None)
        self.output0 = Conv2d(cin, output_channels, 3, 1, weight_initializer, output_func, False)
        self.fused_tail = os.environ.get('RIDERS_FUSED_TAIL', '1') != '0'      # ops.decoder_tail where TF32 is allowed
        # which kernel runs the fused tail: 'tcgen05' (decoder_tail_tc.cu: tcgen05.mma / TMEM, fp16 operands; a value outside the
        # fp16 range raises a device flag and the caller re-runs the mma.sync kernel) or 'mma_sync' (decoder_tail.cu, TF32)
        self.tail_impl = os.environ.get('RIDERS_TAIL_IMPL', 'tcgen05')
        self.tail_fp16_map = os.environ.get('RIDERS_TAIL_FP16_MAP', '1') != '0'   # tcgen05 phase convolution writes the tail's fp16 operand map
        self.tail_upconv = os.environ.get('RIDERS_TAIL_UPCONV', '1') != '0'       # csrc/upconv_tc.cu (else the TMA-fed phase convolution)
        self.tail_upconv_l1 = os.environ.get('RIDERS_UPCONV_L1', '1') != '0'      # the same kernel for the 64 -> 32 block in front of it
        self.force_tail_impl = None                     # set by forward_output_batched for its re-run after a range violation
        self.defer_tail_check = False                   # True: the range flag is read at the caller's own sync point
        self.subconv_tc = os.environ.get('RIDERS_SUBCONV_TC', '0') != '0'      # ops.subpixel_conv_tc instead of cuDNN's 2x2 convolution: opt-in (both sit at the HBM bound of this write-heavy shape, cuDNN 10-15 % ahead)
        self.fused_sub = os.environ.get('RIDERS_FUSED_SUBCONV', '0') != '0'    # ops.subpixel_conv_act in front of it: opt-in, measured no faster than bias_act + cuDNN

    def _tail_ok(self, last, y, yp, shape):
        """Can ``ops.decoder_tail`` take over from the last block's sub-pixel convolution on?  (inference fast path, exact 2x,
        16 -> 16 -> 1 channels, LeakyReLU, BatchNorm folded, no skip)"""
        head, up, cv = self.output0.conv, last.deconv.conv, last.conv
        return (shape is not None and yp is None and last._w4 is not None and last.skip_channels == 0
                and tuple(shape[-2:]) == (2 * y.shape[2], 2 * y.shape[3]) and up.fast_ok(y)
                and up._fused_slope == _LEAK and cv._fused_slope == _LEAK and not cv.use_batch_norm
                and up.conv.out_channels == 16 and tuple(cv.conv.weight.shape) == (16, 16, 3, 3)
                and cv.conv.stride == (1, 1) and cv.conv.padding == (1, 1)
                and self.output0._act is None and not self.output0.use_batch_norm and head.bias is None
                and tuple(head.weight.shape) == (1, 16, 3, 3) and head.stride == (1, 1) and head.padding == (1, 1))

    def _sub_ok(self, last, shape):
        """May the block in front of the last one hand over its raw convolution output (own TF32 phase-convolution kernel
        with that block's bias + LeakyReLU fused in)?  Only on the fused-tail path with the 2x2-tap 32 -> 4 x 16 kernels."""
        return (self.fused_sub and self.fused_tail and torch.backends.cudnn.allow_tf32 and not torch.is_grad_enabled()
                and shape is not None and last.subpixel_taps == 2 and last._w2 is not None
                and tuple(last._w2.shape) == (64, 32, 2, 2) and last.skip_channels == 0)

    def _upconv_ok(self, last, shape):
        """May the block in front of the last one hand over its raw convolution output to ``ops.upconv_tc_half`` (the tcgen05
        up-convolution kernel applies that block's bias + LeakyReLU while it stages its operand)?"""
        return ((self.force_tail_impl or self.tail_impl) == 'tcgen05' and self.tail_fp16_map and self.tail_upconv
                and self.fused_tail and torch.backends.cudnn.allow_tf32 and not torch.is_grad_enabled() and shape is not None
                and last.subpixel_taps == 2 and last._w2 is not None and tuple(last._w2.shape) == (64, 32, 2, 2)
                and last.skip_channels == 0)

    def skip_lead_channels(self, n_skips):
        """Output channels of the up-convolution in front of each skip (skips[i] feeds block ``_order[n_skips-1-i]``,
        see ``forward``): the number of channels a placed skip must leave free in its concat buffer."""
        return [getattr(self, self._order[n_skips - 1 - i]).deconv.conv.conv.out_channels for i in range(n_skips)]

    def forward(self, x, skips, shape=None, head_sigmoid=False):
        """``head_sigmoid`` (inference fast path only; ignored otherwise - check ``self.last_head_sigmoid``): let the fused
        head kernel write torch.sigmoid(logits), so that the reduce kernel streams confidences."""
        self.last_head_sigmoid = False
        self.last_used_tail = False                    # True when the TF32 fused tail ran (tests / bench label the mode by it)
        n = len(skips) - 1
        y, yp = x, None
        raw_bias = raw_slope = None                    # set when y is a RAW convolution output (see want_raw below)
        for name in self._order:
            blk = getattr(self, name)
            blk.upconv_tc = (self.tail_upconv_l1 and self.fused_tail and (self.force_tail_impl or self.tail_impl) == 'tcgen05')
            blk.upconv_defer = self.defer_tail_check
        for i, name in enumerate(self._order[:-1]):
            # target of the NEXT block: when it is 2n+1 along one axis, this block's bias + activation pass writes the
            # replicate-padded operand of that block's sub-pixel convolution directly (DecoderBlock.forward, pad_out)
            nxt = getattr(self, self._order[i + 1])
            tgt = skips[n - 1].shape[2:4] if n >= 1 else (tuple(shape[-2:]) if shape is not None else None)
            cur = skips[n].shape[2:4]
            pad_out = None
            if tgt is not None and nxt._w4 is not None and not torch.is_grad_enabled():
                sy, sx = int(tgt[0] == 2 * cur[0] + 1), int(tgt[1] == 2 * cur[1] + 1)
                if sy + sx == 1 and tgt[0] == 2 * cur[0] + sy and tgt[1] == 2 * cur[1] + sx:
                    pad_out = (sy, sx)
            # the block in front of the fused tail hands over its RAW convolution output: its bias + LeakyReLU are applied
            # by the tail's own phase-convolution kernel while it stages its operand (ops.subpixel_conv_act)
            want_raw = (i == len(self._order) - 2 and pad_out is None and getattr(self, name).conv._fused_slope is not None
                        and (self._sub_ok(getattr(self, self._order[-1]), shape)
                             or self._upconv_ok(getattr(self, self._order[-1]), shape)))
            # ... and so does the block in front of a block whose up-convolution runs as ops.upconv_tc_f32
            if (not want_raw and i < len(self._order) - 2 and pad_out is None and tgt is not None
                    and getattr(self, name).conv._fused_slope is not None and 0.0 <= getattr(self, name).conv._fused_slope <= 1.0
                    and nxt.upconv_tc_ok((y.shape[0], getattr(self, name).conv.conv.out_channels, cur[0], cur[1]), tgt)):
                want_raw = True
            x_act = None if raw_slope is None else (raw_bias, raw_slope)
            out = getattr(self, name)(y, skips[n], xp=yp, pad_out=pad_out, raw=want_raw, x_act=x_act)
            raw_bias = raw_slope = None
            if want_raw:
                y, raw_bias, raw_slope = out
                yp = None
            else:
                y, yp = out if isinstance(out, tuple) else (out, None)
            n -= 1
        last = getattr(self, self._order[-1])
        head = self.output0.conv
        if raw_slope is not None and not (n != 0 and self._tail_ok(last, y, yp, shape) and raw_slope is not None
                                          and 0.0 <= raw_slope <= 1.0 and y.shape[1] == 32):
            if ops._is_channels_last(y):                                                   # not fusable after all
                y = ops.bias_act_(y, raw_bias, raw_slope) if (raw_bias is not None or raw_slope != 1.0) else y
            else:
                y = F.leaky_relu(y if raw_bias is None else y + raw_bias.view(1, -1, 1, 1), raw_slope)
            raw_bias = raw_slope = None
        if (n != 0 and self.fused_tail and torch.backends.cudnn.allow_tf32 and self._tail_ok(last, y, yp, shape)):
            # TF32 flow (torch's default): the whole tail - depth-to-space + bias + LeakyReLU, the block's 16 -> 16
            # convolution, bias + LeakyReLU, the 16 -> 1 head - in ONE kernel fed by the sub-pixel phase convolution
            taps = 2 if (last.subpixel_taps == 2 and last._w2 is not None) else 3
            impl = self.force_tail_impl or self.tail_impl
            if (impl == 'tcgen05' and self.tail_fp16_map and taps == 2 and y.shape[1] == 32
                    and tuple(last._w2.shape) == (64, 32, 2, 2) and y.shape[3] <= 54 and ops._is_channels_last(y)
                    and (raw_slope is None or self.tail_upconv)):
                # the up-convolution in one tcgen05 kernel that writes the tail's fp16 operand map, then the tcgen05 tail.
                # 'upconv': operands staged once, the previous block's bias + LeakyReLU applied on the way (y may be raw);
                # 'tma': the round-2 TMA-fed TF32 phase convolution with the finishing epilogue (RIDERS_TAIL_UPCONV=0)
                self.last_head_sigmoid = bool(head_sigmoid)
                self.last_used_tail = True
                self.last_tail_impl = 'tcgen05+fp16map'
                return [ops.decoder_upconv_tail_tc(y, last._w2, last.deconv.conv.conv.bias, last.conv.conv.weight,
                                                   last.conv.conv.bias, head.weight, _LEAK, sigmoid=head_sigmoid,
                                                   defer=self.defer_tail_check, bias_p=raw_bias,
                                                   slope_p=1.0 if raw_slope is None else raw_slope,
                                                   producer='upconv' if self.tail_upconv else 'tma')]
            if raw_slope is not None:
                y4 = ops.subpixel_conv_act(y, raw_bias, raw_slope, last._w2)          # previous block's bias + LeakyReLU fused in
            elif (self.subconv_tc and taps == 2 and tuple(last._w2.shape) == (64, 32, 2, 2) and y.shape[3] <= 54
                  and ops._is_channels_last(y)):
                y4 = ops.subpixel_conv_tc(y, last._w2)                                # TMA -> tcgen05 -> TMEM phase convolution
            else:
                y4 = F.conv2d(y, last._w2 if taps == 2 else last._w4, None, 1, 1)
            self.last_head_sigmoid = bool(head_sigmoid)
            self.last_used_tail = True
            self.last_tail_impl = impl
            if impl == 'tcgen05':
                return [ops.decoder_tail_tc(y4, last.deconv.conv.conv.bias, last.conv.conv.weight, last.conv.conv.bias,
                                            head.weight, _LEAK, taps=taps, sigmoid=head_sigmoid, defer=self.defer_tail_check)]
            return [ops.decoder_tail(y4, last.deconv.conv.conv.bias, last.conv.conv.weight, last.conv.conv.bias, head.weight,
                                     _LEAK, taps=taps, sigmoid=head_sigmoid)]
        if (n != 0 and last.skip_channels == 0 and last.conv.fast_ok(y) and self.output0._act is None
                and not self.output0.use_batch_norm and head.bias is None and head.kernel_size == (3, 3)
                and head.out_channels == 1 and head.in_channels == 16 and head.stride == (1, 1) and head.padding == (1, 1)):
            # inference fast path: last conv raw -> [bias + LeakyReLU + 3x3 head] in one kernel -> logits
            y_raw, bias, slope = last(y, shape=tuple(shape[-2:]), raw=True, xp=yp)
            if ops._is_channels_last(y_raw):
                self.last_head_sigmoid = bool(head_sigmoid)
                return [ops.conv3x3_head(y_raw, bias, slope, head.weight, sigmoid=head_sigmoid)]
            y = ops.bias_act_(y_raw, bias, slope) if bias is not None or slope != 1.0 else y_raw
            return [self.output0(y)]
        y = last(y, skips[n], xp=yp) if n == 0 else last(y, shape=tuple(shape[-2:]), xp=yp)
        return [self.output0(y)]
