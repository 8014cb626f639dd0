"""LoFTR linear-attention modules, host side (mirror of the reference's
``RCNet/linear_attention.py``; same class names, constructor arguments, parameter names
``q_proj k_proj v_proj merge mlp.0 mlp.2 norm1 norm2`` and call signatures, so reference
checkpoints load and ``networks.py`` can import it unchanged).

Forward passes run on the sm_100a kernels:
  LinearAttention.forward          -> riders_linear_attention_fwd   (linear_attention.py:18-45)
  LoFTREncoderLayer.forward        -> riders_loftr_layer_fwd        (linear_attention.py:112-135)
  LocalFeatureTransformer.forward  -> riders_loftr_transformer_fwd  (linear_attention.py:159-184)
``FullAttention`` (linear_attention.py:49-80) is never constructed by RC-Net (attention='linear', networks.py:378); it is
provided as plain torch ops for API completeness.

fp16-range guard: the tcgen05 path computes in fp16 hi/lo pairs (|value| <= 65504, a limit the fp32 reference does not have).
The kernel flags a violation and the call is recomputed on the fp32 SIMT kernels - immediately here (one small D2H per call),
or at the caller's own sync point when ``defer_overflow_check`` is set (``forward_output_batched``).
"""
import copy
import os

import torch
import torch.nn as nn

from .. import ops


def default_impl():
    """RIDERS_LOFTR_IMPL=simt|tcgen05 selects the kernel family (default: tcgen05)."""
    name = os.environ.get("RIDERS_LOFTR_IMPL", "tcgen05").lower()
    if name not in ("simt", "tcgen05"):
        raise ValueError("RIDERS_LOFTR_IMPL must be 'simt' or 'tcgen05'")
    return ops.LOFTR_SIMT if name == "simt" else ops.LOFTR_TCGEN05


class LinearAttention(nn.Module):
    def __init__(self, eps=1e-6):
        super().__init__()
        self.eps = eps

    def forward(self, queries, keys, values, q_mask=None, kv_mask=None):
        """queries [N,L,H,D], keys/values [N,S,H,D], masks [N,L]/[N,S] -> [N,L,H,D] contiguous."""
        return ops.linear_attention(queries, keys, values, q_mask, kv_mask, self.eps)


class FullAttention(nn.Module):
    """linear_attention.py:49-80: softmax attention with the 1/sqrt(D) temperature (never constructed by RC-Net; torch ops)."""

    def __init__(self, use_dropout=False, attention_dropout=0.1):
        super().__init__()
        self.use_dropout = use_dropout
        self.dropout = nn.Dropout(attention_dropout)

    def forward(self, queries, keys, values, q_mask=None, kv_mask=None):
        QK = torch.einsum("nlhd,nshd->nlsh", queries, keys)
        if kv_mask is not None:
            QK.masked_fill_(~(q_mask[:, :, None, None] * kv_mask[:, None, :, None]).bool(), float('-inf'))
        softmax_temp = 1. / queries.size(3) ** .5
        A = torch.softmax(softmax_temp * QK, dim=2)
        if self.use_dropout:
            A = self.dropout(A)
        return torch.einsum("nlsh,nshd->nlhd", A, values).contiguous()


class LoFTREncoderLayer(nn.Module):
    def __init__(self, d_model, nhead, attention='linear'):
        super().__init__()
        self.dim = d_model // nhead
        self.nhead = nhead
        self.q_proj = nn.Linear(d_model, d_model, bias=False)
        self.k_proj = nn.Linear(d_model, d_model, bias=False)
        self.v_proj = nn.Linear(d_model, d_model, bias=False)
        self.attention = LinearAttention() if attention == 'linear' else FullAttention()
        self.merge = nn.Linear(d_model, d_model, bias=False)
        self.mlp = nn.Sequential(
            nn.Linear(d_model * 2, d_model * 2, bias=False),
            nn.ReLU(True),
            nn.Linear(d_model * 2, d_model, bias=False),
        )
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)

    def kernel_params(self):
        return dict(self.named_parameters())

    def _supports(self, impl, x, source):
        # both families are CUDA kernels of this library; the tcgen05 layer is specialised
        if x.shape[1] != source.shape[1]:
            return impl == ops.LOFTR_SIMT
        return ops.loftr_impl_available(impl, x.shape[2], self.nhead, x.shape[1])

    def _forward_autograd(self, x, source):
        """Training path (rcnet_main.py:342-358 needs gradients): the hand-written forward / backward kernels behind
        ``ops.LoftrLayerFunction`` (riders_loftr_layer_fwd_train / riders_loftr_layer_bwd).  Under bf16 autocast the
        tokens are bfloat16, the parameters stay the float32 master weights."""
        if torch.is_autocast_enabled() and torch.get_autocast_gpu_dtype() == torch.bfloat16:
            same = source is x
            x = x.to(torch.bfloat16)
            source = x if same else source.to(torch.bfloat16)
        elif x.dtype != source.dtype:
            source = source.to(x.dtype)
        return ops.loftr_layer_autograd(x, source, self.kernel_params(), self.nhead, self.attention.eps, self.norm1.eps)

    def forward(self, x, source, x_mask=None, source_mask=None):
        """x [N,L,C], source [N,S,C] -> [N,L,C]."""
        if torch.is_grad_enabled() and x_mask is None and source_mask is None and isinstance(self.attention, LinearAttention) and \
                (x.requires_grad or source.requires_grad or self.q_proj.weight.requires_grad):
            return self._forward_autograd(x, source)
        if x_mask is not None or source_mask is not None or not isinstance(self.attention, LinearAttention):
            # masked call / FullAttention (never issued by RC-Net, networks.py:444): compose from the kernels
            bs = x.size(0)
            q = self.q_proj(x).view(bs, -1, self.nhead, self.dim)
            k = self.k_proj(source).view(bs, -1, self.nhead, self.dim)
            v = self.v_proj(source).view(bs, -1, self.nhead, self.dim)
            m = self.attention(q, k, v, q_mask=x_mask, kv_mask=source_mask)
            m = self.norm1(self.merge(m.view(bs, -1, self.nhead * self.dim)))
            m = self.norm2(self.mlp(torch.cat([x, m], dim=2)))
            return x + m
        impl = default_impl()
        if not self._supports(impl, x, source):
            impl = ops.LOFTR_SIMT
        return ops.loftr_layer(x, source, self.kernel_params(), self.nhead, impl,
                               self.attention.eps, self.norm1.eps)


class LocalFeatureTransformer(nn.Module):
    """A Local Feature Transformer (LoFTR) module: ['self','cross'] * n_layers."""

    def __init__(self, type, n_layers=1, d_model=256, nhead=8, attention='linear'):
        super().__init__()
        self.d_model = d_model
        self.nhead = nhead
        self.layer_names = type * n_layers
        self.attention = attention
        encoder_layer = LoFTREncoderLayer(self.d_model, self.nhead, self.attention)
        self.layers = nn.ModuleList([copy.deepcopy(encoder_layer) for _ in range(len(self.layer_names))])
        self._reset_parameters()

    def _reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def forward(self, feat0, feat1, mask0=None, mask1=None):
        assert self.d_model == feat0.size(2), "the feature number of src and transformer must be equal"
        if mask0 is not None or mask1 is not None or self.attention != 'linear':
            for layer, name in zip(self.layers, self.layer_names):
                if name == 'self':
                    feat0 = layer(feat0, feat0, mask0, mask0)
                    feat1 = layer(feat1, feat1, mask1, mask1)
                elif name == 'cross':
                    feat0 = layer(feat0, feat1, mask0, mask1)
                    feat1 = layer(feat1, feat0, mask1, mask0)
                else:
                    raise KeyError
            return feat0, feat1
        if torch.is_grad_enabled() and (feat0.requires_grad or feat1.requires_grad or
                                        self.layers[0].q_proj.weight.requires_grad):
            for layer, name in zip(self.layers, self.layer_names):            # training: differentiable torch ops
                if name == 'self':
                    feat0 = layer(feat0, feat0)
                    feat1 = layer(feat1, feat1)
                elif name == 'cross':
                    feat0 = layer(feat0, feat1)
                    feat1 = layer(feat1, feat0)
                else:
                    raise KeyError
            return feat0, feat1
        impl = getattr(self, "force_impl", None)
        impl = default_impl() if impl is None else impl
        if not self.layers[0]._supports(impl, feat0, feat1):
            impl = ops.LOFTR_SIMT
        return ops.loftr_transformer(feat0, feat1, [l.kernel_params() for l in self.layers],
                                     self.layer_names, self.nhead, impl,
                                     self.layers[0].attention.eps, self.layers[0].norm1.eps,
                                     defer=bool(getattr(self, "defer_overflow_check", False)))
