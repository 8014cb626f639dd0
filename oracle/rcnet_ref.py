"""Oracle for SURVEY 8(a) rows A0/A5 end to end: RC-Net forward + ``forward_output`` on the CPU.

Restates, with CPU torch / torchvision / NumPy only:
  * ``RCNetEncoder.forward``  RCNet/networks.py:389-451 (roi_pool x5, point MLP, token
    reshape, LoFTR transformer, concat)
  * ``RCNetModel.forward``    RCNet/rcnet_model.py:101-123
  * ``forward_output``        RCNet/rcnet_main.py:435-487 (literal canvas/cat/max formulation
    for small K, and the streaming restatement of oracle.confidence_reduce_ref)
The CNN backbone / point MLP / decoder are taken from the model object as plain torch modules
(they are not part of the kernels under test); everything the CUDA kernels replace is recomputed
here with library CPU ops.  Also used, timed on the host cores, as the ``cpu_baseline`` /
``--impl reference`` arm of bench.py (kind "port").  Test infrastructure only.
"""
import numpy as np
import torch
import torchvision

from .confidence_reduce_ref import confidence_reduce_ref


def _unwrap(m):
    return m.module if hasattr(m, "module") else m


def loftr_layer_torch(layer, x, source):
    """RCNet/linear_attention.py:112-135 + 18-45 with torch CPU ops (fp32, as shipped)."""
    bs = x.size(0)
    H, D = layer.nhead, layer.dim
    q = torch.nn.functional.linear(x, layer.q_proj.weight).view(bs, -1, H, D)
    k = torch.nn.functional.linear(source, layer.k_proj.weight).view(bs, -1, H, D)
    v = torch.nn.functional.linear(source, layer.v_proj.weight).view(bs, -1, H, D)
    Q = torch.nn.functional.elu(q) + 1
    K = torch.nn.functional.elu(k) + 1
    S = v.size(1)
    v = v / S
    KV = torch.einsum("nshd,nshv->nhdv", K, v)
    Z = 1 / (torch.einsum("nlhd,nhd->nlh", Q, K.sum(dim=1)) + 1e-6)
    msg = (torch.einsum("nlhd,nhdv,nlh->nlhv", Q, KV, Z) * S).contiguous()
    msg = torch.nn.functional.linear(msg.view(bs, -1, H * D), layer.merge.weight)
    msg = torch.nn.functional.layer_norm(msg, (H * D,), layer.norm1.weight, layer.norm1.bias, layer.norm1.eps)
    h = torch.relu(torch.nn.functional.linear(torch.cat([x, msg], dim=2), layer.mlp[0].weight))
    msg = torch.nn.functional.linear(h, layer.mlp[2].weight)
    msg = torch.nn.functional.layer_norm(msg, (H * D,), layer.norm2.weight, layer.norm2.bias, layer.norm2.eps)
    return x + msg


def transformer_torch(tf, feat0, feat1):
    """RCNet/linear_attention.py:159-184."""
    for layer, name in zip(tf.layers, tf.layer_names):
        if name == "self":
            feat0 = loftr_layer_torch(layer, feat0, feat0)
            feat1 = loftr_layer_torch(layer, feat1, feat1)
        elif name == "cross":
            feat0 = loftr_layer_torch(layer, feat0, feat1)
            feat1 = loftr_layer_torch(layer, feat1, feat0)
        else:
            raise KeyError(name)
    return feat0, feat1


@torch.no_grad()
def encoder_forward_ref(encoder, image, points, b_boxes):
    enc = _unwrap(encoder)
    ph, pw = enc.input_patch_size_image[-2], enc.input_patch_size_image[-1]
    latent_h, latent_w = int(ph // 32.0), int(pw // 32.0)
    skip_scales = [1 / 2.0, 1 / 4.0, 1 / 8.0, 1 / 16.0, 1 / 32.0, 1 / 64.0, 1 / 128.0]
    latent_image, skips_image = enc.encoder_image(image)
    boxes = [torch.as_tensor(b, dtype=torch.float32).to(image.device) for b in b_boxes]
    latent_pooled = torchvision.ops.roi_pool(latent_image, boxes, spatial_scale=1 / 32.0,
                                             output_size=(latent_h, latent_w))
    skips_pooled = [torchvision.ops.roi_pool(s, boxes, spatial_scale=skip_scales[i],
                                             output_size=(int(ph * skip_scales[i]), int(pw * skip_scales[i])))
                    for i, s in enumerate(skips_image)]
    c = enc.n_neuron_latent_depth
    latent_depth = enc.encoder_depth(points).view(points.shape[0], c, -1, latent_w)
    tok_d = latent_depth.view(latent_depth.shape[0], c, -1).permute(0, 2, 1)
    tok_i = latent_pooled.view(latent_pooled.shape[0], latent_pooled.shape[1], -1).permute(0, 2, 1)
    tok_d, tok_i = transformer_torch(enc.attention, tok_d, tok_i)
    depth_tf = tok_d.permute(0, 2, 1).reshape(latent_depth.shape)
    image_tf = tok_i.permute(0, 2, 1).reshape(latent_pooled.shape)
    return torch.cat([image_tf, depth_tf], dim=1), skips_pooled


@torch.no_grad()
def model_forward_ref(model, image, points, b_boxes, return_logits=True):
    latent, skips = encoder_forward_ref(model.encoder, image, points, b_boxes)
    logits = _unwrap(model.decoder)(x=latent, skips=skips, shape=model.input_patch_size_image)[-1]
    return logits if return_logits else torch.sigmoid(logits)


@torch.no_grad()
def forward_output_ref(model, image, radar_points, bounding_boxes_list, response_thr=0.5, literal=False, want_conf=False):
    """CPU restatement of rcnet_main.py:435-487.  Returns depth (1,H,W), response (1,H,W), select (1,H,W)."""
    ph, pw = model.input_patch_size_image
    px, py = pw // 2, ph // 2
    H, W = image.shape[-2], image.shape[-1]
    padded = torch.nn.functional.pad(image, (px, px, py, py), mode="replicate")
    if radar_points.dim() == 3:
        radar_points = radar_points.squeeze(0)
    conf = model_forward_ref(model, padded, radar_points, bounding_boxes_list, return_logits=False)
    if literal:
        tiles = []
        for crop, p in zip(conf, radar_points):
            canvas = torch.zeros([1, padded.shape[-2], padded.shape[-1]], device=conf.device)
            crop = torch.where(crop < response_thr, torch.zeros_like(crop), crop)
            x, y = int(p[0]), int(p[1])
            canvas[:, y - py:y + py, x - px:x + px] = crop
            tiles.append(canvas)
        tiles = torch.cat(tiles, dim=0)[:, py:-py, px:-px]
        response, select = torch.max(tiles, dim=0, keepdim=True)
        depth = torch.sum(tiles * radar_points[:, 2].unsqueeze(1).unsqueeze(1), dim=0, keepdim=True) / \
            torch.sum(tiles, dim=0, keepdim=True)
        depth = torch.where(response == 0, torch.zeros_like(depth), depth)
        if want_conf:
            return depth, response, select, conf[:, 0].cpu().numpy()
        return depth, response, select
    d, r, s = confidence_reduce_ref(conf[:, 0].numpy(), radar_points.numpy(), H, W, response_thr)
    if want_conf:
        return torch.from_numpy(d)[None], torch.from_numpy(r)[None], torch.from_numpy(s)[None], conf[:, 0].numpy()
    return torch.from_numpy(d)[None], torch.from_numpy(r)[None], torch.from_numpy(s)[None]
