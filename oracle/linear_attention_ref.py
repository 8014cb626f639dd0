"""Oracle for SURVEY 8(a) rows A2-A4: LoFTR linear-attention fusion.

Follows RCNet/linear_attention.py of the reference:
  * A2 ``LinearAttention.forward``           lines 18-45  (elu+1 feature map, eps 1e-6 line 13)
  * A3 ``LoFTREncoderLayer.forward``         lines 112-135 (bias-free projections lines 95-106,
                                              LayerNorm eps 1e-5 lines 109-110)
  * A4 ``LocalFeatureTransformer.forward``   lines 159-184 (['self','cross']*4, networks.py:378)

NumPy restatement parameterised by dtype (float32 = as shipped, float64 =
ground truth used to budget tolerances).  Weights are taken from a plain
dict keyed like the reference state_dict (``layers.{i}.q_proj.weight`` ...).
Test infrastructure only.
"""
import numpy as np


def elu_plus_one(x):
    # linear_attention.py:7-8 : elu(x) + 1
    return np.where(x > 0, x + 1, np.exp(np.minimum(x, 0)))  # elu(x)+1 = exp(x) for x<=0


def linear_attention_ref(q, k, v, q_mask=None, kv_mask=None, eps=1e-6):
    """q (N,L,H,D), k/v (N,S,H,D) -> (N,L,H,D); linear_attention.py:18-45."""
    dt = q.dtype
    Q = elu_plus_one(q).astype(dt)
    K = elu_plus_one(k).astype(dt)
    if q_mask is not None:
        Q = Q * q_mask[:, :, None, None].astype(dt)
    if kv_mask is not None:
        K = K * kv_mask[:, :, None, None].astype(dt)
        v = v * kv_mask[:, :, None, None].astype(dt)
    S = v.shape[1]
    v = v / dt.type(S)
    KV = np.einsum("nshd,nshv->nhdv", K, v)
    Z = 1 / (np.einsum("nlhd,nhd->nlh", Q, K.sum(axis=1)) + dt.type(eps))
    out = np.einsum("nlhd,nhdv,nlh->nlhv", Q, KV, Z) * dt.type(S)
    return np.ascontiguousarray(out.astype(dt))


def layer_norm(x, g, b, eps=1e-5):
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + x.dtype.type(eps)) * g + b


def loftr_layer_ref(x, src, w, nhead=8, x_mask=None, source_mask=None):
    """w: dict with q_proj,k_proj,v_proj,merge,mlp.0,mlp.2 (.weight) and norm1/norm2 (.weight/.bias)."""
    dt = x.dtype
    W = {k: np.asarray(val, dtype=dt) for k, val in w.items()}
    N, L, C = x.shape
    D = C // nhead
    q = (x @ W["q_proj.weight"].T).reshape(N, -1, nhead, D)
    k = (src @ W["k_proj.weight"].T).reshape(N, -1, nhead, D)
    v = (src @ W["v_proj.weight"].T).reshape(N, -1, nhead, D)
    msg = linear_attention_ref(q, k, v, x_mask, source_mask).reshape(N, -1, C)
    msg = msg @ W["merge.weight"].T
    msg = layer_norm(msg, W["norm1.weight"], W["norm1.bias"])
    h = np.concatenate([x, msg], axis=2) @ W["mlp.0.weight"].T
    h = np.maximum(h, 0)
    msg = h @ W["mlp.2.weight"].T
    msg = layer_norm(msg, W["norm2.weight"], W["norm2.bias"])
    return (x + msg).astype(dt)


def split_layer_weights(state, n_layers=8, prefix="layers."):
    out = []
    for i in range(n_layers):
        p = f"{prefix}{i}."
        out.append({k[len(p):]: np.asarray(v) for k, v in state.items() if k.startswith(p)})
    return out


def transformer_ref(feat0, feat1, state, layer_names=("self", "cross") * 4, nhead=8, dtype=np.float32):
    """linear_attention.py:159-184.  state: reference-style state_dict (numpy values)."""
    f0 = np.asarray(feat0, dtype=dtype)
    f1 = np.asarray(feat1, dtype=dtype)
    layers = split_layer_weights(state, len(layer_names))
    for w, name in zip(layers, layer_names):
        if name == "self":
            f0 = loftr_layer_ref(f0, f0, w, nhead)
            f1 = loftr_layer_ref(f1, f1, w, nhead)
        elif name == "cross":
            f0 = loftr_layer_ref(f0, f1, w, nhead)
            f1 = loftr_layer_ref(f1, f0, w, nhead)
        else:
            raise KeyError(name)
    return f0, f1
