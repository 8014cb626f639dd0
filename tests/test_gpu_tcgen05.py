"""tcgen05 / TMEM path of the LoFTR fusion (A3/A4): device self-test of the operand staging +
descriptors, then the fused layer and the 8-layer transformer against the reference-generated
golden vectors (tests/golden/loftr.npz, case 'rcnet': d_model 128, 8 heads, 21 tokens)."""
import ctypes as C

import numpy as np
import pytest
import torch

from riders_b200 import _lib, ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _gemm(a, b, terms):
    K = a.shape[1]
    d = torch.empty((128, 128), dtype=torch.float32, device=DEV)
    ta, tb = cu(a), cu(b)                      # keep both alive: the caching allocator would reuse a freed block
    rc = _lib.lib().riders_debug_tc_gemm(C.c_void_p(ta.data_ptr()), C.c_void_p(tb.data_ptr()),
                                         C.c_void_p(d.data_ptr()), K, terms,
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "riders_debug_tc_gemm")
    torch.cuda.synchronize()
    return d.cpu().numpy()


@pytest.mark.parametrize("K", [64, 128])
def test_tc_gemm_selftest(K):
    rng = np.random.default_rng(K)
    a = rng.standard_normal((128, K)).astype(np.float32)
    b = rng.standard_normal((128, K)).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    scale = np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64).T
    # small integers are exact in fp16: any layout / descriptor mistake shows up as a wrong integer
    ai = rng.integers(-8, 9, (128, K)).astype(np.float32)
    bi = rng.integers(-8, 9, (128, K)).astype(np.float32)
    assert np.array_equal(_gemm(ai, bi, 1), ai @ bi.T)
    one = _gemm(a, b, 1)
    assert np.max(np.abs(one - ref) / scale) < 2e-3            # plain fp16: 2^-11 per operand
    three = _gemm(a, b, 3)
    assert np.max(np.abs(three - ref) / scale) < 2e-6           # hi/lo split, 3 MMAs: fp32-class
    assert np.max(np.abs(three - ref) / scale) < 0.01 * np.max(np.abs(one - ref) / scale)


def _layer_params(g, i):
    pre = "layers.%d." % i
    return {k[len(pre):]: cu(v) for k, v in g.items() if k.startswith(pre)}


def test_loftr_tcgen05_layer_and_transformer_match_reference(golden):
    assert ops.loftr_impl_available(ops.LOFTR_TCGEN05, 128, 8, 21)
    g = golden("loftr.npz").case("rcnet")
    n_layers = int(g["n_layers"])
    f0, f1 = cu(g["f0"]), cu(g["f1"])
    # self layer (x == source) and cross layer (x != source) against the SIMT kernels and the golden
    one = ops.loftr_layer(f0, f0, _layer_params(g, 0), impl=ops.LOFTR_TCGEN05)
    ref_self = ops.loftr_layer(f0, f0, _layer_params(g, 0), impl=ops.LOFTR_SIMT)
    np.testing.assert_allclose(one.cpu().numpy(), ref_self.cpu().numpy(), rtol=1e-4, atol=2e-5)
    cross = ops.loftr_layer(f0, f1, _layer_params(g, 0), impl=ops.LOFTR_TCGEN05)
    np.testing.assert_allclose(cross.cpu().numpy(), g["layer0"], rtol=1e-4, atol=2e-5)
    o0, o1 = ops.loftr_transformer(f0, f1, [_layer_params(g, i) for i in range(2 * n_layers)],
                                   ["self", "cross"] * n_layers, impl=ops.LOFTR_TCGEN05)
    np.testing.assert_allclose(o0.cpu().numpy(), g["o0"], rtol=1e-4, atol=5e-5)      # 1e-4 relative (north_star)
    np.testing.assert_allclose(o1.cpu().numpy(), g["o1"], rtol=1e-4, atol=5e-5)


def _random_layer(seed):
    rng = np.random.default_rng(seed)
    g = torch.Generator(device="cpu").manual_seed(seed)
    params = {}
    for name, shape in (("q_proj.weight", (128, 128)), ("k_proj.weight", (128, 128)), ("v_proj.weight", (128, 128)),
                        ("merge.weight", (128, 128)), ("mlp.0.weight", (256, 256)), ("mlp.2.weight", (128, 256))):
        w = torch.empty(shape)
        torch.nn.init.xavier_uniform_(w, generator=g)
        params[name] = w.to(DEV)
    for name in ("norm1", "norm2"):
        params[name + ".weight"] = cu(rng.uniform(0.5, 1.5, 128).astype(np.float32))
        params[name + ".bias"] = cu(rng.uniform(-0.2, 0.2, 128).astype(np.float32))
    return params


def test_loftr_tcgen05_fp16_range_guard():
    """The tcgen05 path works in fp16 hi/lo pairs (|value| <= 65504); the fp32 reference has no such limit.  Tokens of
    magnitude 1e5 must not yield silent Inf / NaN: the kernel raises its overflow flag and the call is recomputed on the
    fp32 SIMT kernels (immediately, or at the caller's sync point in the deferred form)."""
    params = _random_layer(3)
    rng = np.random.default_rng(3)
    big = cu((rng.standard_normal((13, 21, 128)) * 1e5).astype(np.float32))
    ok = cu(rng.standard_normal((13, 21, 128)).astype(np.float32))
    ref = ops.loftr_layer(big, big, params, impl=ops.LOFTR_SIMT)
    got = ops.loftr_layer(big, big, params, impl=ops.LOFTR_TCGEN05)
    assert torch.isfinite(got).all() and torch.equal(got, ref)
    # one oversized token in an otherwise ordinary batch, cross layer
    mixed = ok.clone()
    mixed[7, 3, 5] = 3e5
    got = ops.loftr_layer(ok, mixed, params, impl=ops.LOFTR_TCGEN05)
    assert torch.equal(got, ops.loftr_layer(ok, mixed, params, impl=ops.LOFTR_SIMT))
    # deferred form: the flag is read later, by the caller
    assert not ops.LoftrOverflow.pending
    ops.loftr_layer(ok, ok, params, impl=ops.LOFTR_TCGEN05, defer=True)
    assert ops.LoftrOverflow.pending and not ops.loftr_overflow_pending()
    ops.loftr_layer(big, big, params, impl=ops.LOFTR_TCGEN05, defer=True)
    ops.loftr_transformer(ok, ok, [params, params], ["self", "cross"], impl=ops.LOFTR_TCGEN05, defer=True)
    assert ops.loftr_overflow_pending() and not ops.LoftrOverflow.pending
    # the whole transformer with oversized tokens
    o0, o1 = ops.loftr_transformer(big, ok, [params, params], ["self", "cross"], impl=ops.LOFTR_TCGEN05)
    r0, r1 = ops.loftr_transformer(big, ok, [params, params], ["self", "cross"], impl=ops.LOFTR_SIMT)
    assert torch.equal(o0, r0) and torch.equal(o1, r1)


def test_loftr_tcgen05_whole_transformer_is_one_launch():
    """All 2 x n_layers layer calls of LocalFeatureTransformer.forward run inside ONE persistent kernel (+ one weight-pack
    launch per layer): a tile's tokens stay with its CTA from the first layer call to the last."""
    params = [_random_layer(10 + i) for i in range(8)]
    rng = np.random.default_rng(5)
    f0 = cu(rng.standard_normal((1000, 21, 128)).astype(np.float32))
    f1 = cu(np.maximum(rng.standard_normal((1000, 21, 128)), 0).astype(np.float32))
    n0 = ops.launch_count()
    o0, o1 = ops.loftr_transformer(f0, f1, params, ["self", "cross"] * 4, impl=ops.LOFTR_TCGEN05)
    assert ops.launch_count() - n0 == 8 + 1
    r0, r1 = ops.loftr_transformer(f0, f1, params, ["self", "cross"] * 4, impl=ops.LOFTR_SIMT)
    np.testing.assert_allclose(o0.cpu().numpy(), r0.cpu().numpy(), rtol=1e-4, atol=5e-5)
    np.testing.assert_allclose(o1.cpu().numpy(), r1.cpu().numpy(), rtol=1e-4, atol=5e-5)
    again = ops.loftr_transformer(f0, f1, params, ["self", "cross"] * 4, impl=ops.LOFTR_TCGEN05)
    assert torch.equal(again[0], o0) and torch.equal(again[1], o1)                       # deterministic


@pytest.mark.parametrize("N", [1, 5, 6, 7, 149 * 6 + 1])
def test_loftr_tcgen05_ragged_tiles_vs_simt(N):
    # tile = 6 points: partial tiles, one tile, many tiles per CTA (> 148 tiles)
    rng = np.random.default_rng(N)
    g = torch.Generator(device="cpu").manual_seed(N)
    params = {}
    for name, shape in (("q_proj.weight", (128, 128)), ("k_proj.weight", (128, 128)), ("v_proj.weight", (128, 128)),
                        ("merge.weight", (128, 128)), ("mlp.0.weight", (256, 256)), ("mlp.2.weight", (128, 256))):
        w = torch.empty(shape)
        torch.nn.init.xavier_uniform_(w, generator=g)
        params[name] = w.to(DEV)
    for name in ("norm1", "norm2"):
        params[name + ".weight"] = cu(rng.uniform(0.5, 1.5, 128).astype(np.float32))
        params[name + ".bias"] = cu(rng.uniform(-0.2, 0.2, 128).astype(np.float32))
    x = cu(rng.standard_normal((N, 21, 128)).astype(np.float32))
    s = cu(rng.standard_normal((N, 21, 128)).astype(np.float32))
    for src in (x, s):
        a = ops.loftr_layer(x, src, params, impl=ops.LOFTR_TCGEN05)
        b = ops.loftr_layer(x, src, params, impl=ops.LOFTR_SIMT)
        np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-4, atol=2e-5)
    again = ops.loftr_layer(x, s, params, impl=ops.LOFTR_TCGEN05)
    assert torch.equal(again, ops.loftr_layer(x, s, params, impl=ops.LOFTR_TCGEN05))      # deterministic
