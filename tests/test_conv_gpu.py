"""Parity of the tcgen05 implicit-GEMM convolution against torch's CPU fp32 conv2d on identically rounded operands.

Floating point: the kernel multiplies 16-bit operands exactly and accumulates in fp32, so against an fp32 reference
on the same rounded inputs the only difference is accumulation order: tolerance 2e-3 * max|y| on fp32 outputs
(K up to 4608), plus one 16-bit rounding step (2^-8 relative for bf16, 2^-11 for fp16) on 16-bit outputs.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _round(a, act_dtype):
    t = torch.from_numpy(np.ascontiguousarray(a, np.float32))
    return t.to(torch.bfloat16 if act_dtype == 1 else torch.float16).float()


def _ref(x, w, bias, residual, dil, pad, stride_h, relu, act_dtype):
    xr = _round(x, act_dtype).permute(0, 3, 1, 2)
    wr = _round(w, act_dtype).permute(0, 3, 1, 2)
    y = F.conv2d(xr, wr, None if bias is None else torch.from_numpy(bias), stride=(stride_h, 1), padding=pad,
                 dilation=dil)
    if residual is not None:
        y = y + _round(residual, act_dtype).permute(0, 3, 1, 2)
    if relu:
        y = torch.relu(y)
    return y.permute(0, 2, 3, 1).contiguous().numpy()


CASES = [
    # name, B, H, W, Cin, Cout, KH, KW, dil, pad, stride_h, relu, residual, out_fp32, act, x_extra, y_extra, n_tile
    ("3x3_c64", 1, 32, 32, 64, 64, 3, 3, (1, 1), (1, 1), 1, False, False, True, 1, 0, 0, 0),
    ("3x3_c128_n256_b2", 2, 20, 24, 128, 256, 3, 3, (1, 1), (1, 1), 1, True, False, True, 1, 0, 0, 0),
    ("1x1_c192", 1, 17, 23, 192, 64, 1, 1, (1, 1), (0, 0), 1, False, False, True, 1, 0, 0, 0),
    ("dil6", 1, 24, 24, 64, 128, 3, 3, (6, 6), (6, 6), 1, False, False, True, 1, 0, 0, 0),
    ("k2_s21_p01", 16, 4, 26, 64, 64, 2, 2, (1, 1), (0, 1), 2, True, False, True, 1, 0, 0, 0),
    ("k2_s1_p0", 16, 2, 27, 64, 64, 2, 2, (1, 1), (0, 0), 1, True, False, True, 1, 0, 0, 0),
    ("sw64_c32", 1, 40, 36, 32, 32, 3, 3, (1, 1), (1, 1), 1, True, False, True, 1, 0, 0, 0),
    ("sw32_c16", 1, 40, 36, 16, 16, 3, 3, (1, 1), (1, 1), 1, True, False, True, 1, 0, 0, 0),
    ("residual_bf16out", 4, 4, 26, 128, 128, 3, 3, (1, 1), (1, 1), 1, True, True, False, 1, 0, 0, 0),
    ("fp16", 2, 16, 50, 64, 128, 3, 3, (1, 1), (1, 1), 1, True, True, False, 0, 0, 0, 0),
    ("head_c37", 8, 1, 26, 256, 37, 1, 1, (1, 1), (0, 0), 1, False, False, True, 1, 0, 0, 0),
    ("views", 1, 30, 28, 64, 64, 3, 3, (1, 1), (1, 1), 1, True, False, False, 1, 64, 32, 0),
    ("deep_k_multi_tile", 64, 4, 26, 512, 512, 3, 3, (1, 1), (1, 1), 1, True, True, False, 1, 0, 0, 0),
    ("n_tile_128_of_512", 2, 16, 16, 64, 512, 3, 3, (1, 1), (1, 1), 1, False, False, True, 1, 0, 0, 128),
    ("big_plane", 1, 160, 120, 64, 64, 3, 3, (1, 1), (1, 1), 1, True, False, False, 1, 0, 0, 0),
    # N = 128 with enough M = 256 tiles for the CTA-pair kernels (cta_group::2): plain, residual, 1x1, odd tile counts
    ("cta2_c64_c128", 4, 160, 128, 64, 128, 3, 3, (1, 1), (1, 1), 1, True, False, False, 0, 0, 0, 0),
    ("cta2_residual", 4, 160, 128, 128, 128, 3, 3, (1, 1), (1, 1), 1, True, True, False, 0, 0, 0, 0),
    ("cta2_1x1_views", 3, 200, 136, 192, 128, 1, 1, (1, 1), (0, 0), 1, True, False, False, 1, 64, 64, 0),
    ("cta2_odd_tiles", 5, 122, 126, 64, 128, 3, 3, (1, 1), (1, 1), 1, False, False, False, 0, 0, 0, 0),
    ("cta2_crnn_layer1", 125, 16, 50, 128, 128, 3, 3, (1, 1), (1, 1), 1, True, True, False, 0, 0, 0, 0),
    # narrow layers (N <= 64): the decoder-tail patterns, odd k-block counts
    ("c32_n64_m256_sw64", 2, 64, 48, 32, 64, 3, 3, (1, 1), (1, 1), 1, True, False, False, 0, 0, 0, 0),
    ("kg_c64_n32_9kb", 2, 64, 48, 64, 32, 3, 3, (1, 1), (1, 1), 1, True, False, False, 0, 0, 0, 0),
    ("kg_c32_n32", 2, 64, 48, 32, 32, 3, 3, (1, 1), (1, 1), 1, True, False, False, 0, 0, 0, 0),
    ("kg_1x1_c192_n64", 2, 64, 48, 192, 64, 1, 1, (1, 1), (0, 0), 1, True, False, False, 0, 0, 0, 0),
    ("kg_c128_n64_18kb", 3, 20, 30, 128, 64, 3, 3, (1, 1), (1, 1), 1, True, True, False, 0, 0, 0, 0),
    ("kg_c32_n16_fp32", 1, 33, 47, 32, 16, 3, 3, (1, 1), (1, 1), 1, False, False, True, 1, 0, 0, 0),
    ("kg_k2_5kb", 16, 4, 26, 80, 64, 2, 2, (1, 1), (0, 1), 2, True, False, True, 1, 0, 0, 0),
    # N = 256 tiles on CTA pairs (LOCR_CONV_CTA2_N256): the 512-channel CRNN / VGG patterns, odd m-tile counts, 1x1
    ("cta2_n256_crnn_c512", 33, 4, 26, 512, 512, 3, 3, (1, 1), (1, 1), 1, True, True, False, 0, 0, 0, 0),
    ("cta2_n256_vgg_c256", 2, 80, 60, 128, 256, 3, 3, (1, 1), (1, 1), 1, True, False, False, 0, 0, 0, 0),
    ("cta2_n256_dil6_c1024", 1, 40, 30, 256, 1024, 3, 3, (6, 6), (6, 6), 1, False, False, False, 0, 0, 0, 0),
    ("cta2_n256_1x1", 3, 40, 30, 1024, 512, 1, 1, (1, 1), (0, 0), 1, True, False, False, 1, 0, 0, 0),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_parity(case):
    from lightly_ocr_b200 import bridge
    (name, B, H, W, Cin, Cout, KH, KW, dil, pad, stride_h, relu, use_res, out_fp32, act, x_extra, y_extra,
     n_tile) = case
    import zlib
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    x = rng.standard_normal((B, H, W, Cin + x_extra)).astype(np.float32)
    w = (rng.standard_normal((Cout, KH, KW, Cin)) / np.sqrt(KH * KW * Cin)).astype(np.float32)
    bias = rng.standard_normal(Cout).astype(np.float32)
    OH = (H + 2 * pad[0] - dil[0] * (KH - 1) - 1) // stride_h + 1
    OW = (W + 2 * pad[1] - dil[1] * (KW - 1) - 1) + 1
    residual = rng.standard_normal((B, OH, OW, Cout)).astype(np.float32) if use_res else None
    y = bridge.test_conv(x, w, bias, residual, dil=dil, pad=pad, stride_h=stride_h, relu=relu, out_fp32=out_fp32,
                         act_dtype=act, n_tile=n_tile, y_pitch=Cout + y_extra)
    ref = _ref(x[..., :Cin], w, bias, residual, dil, pad, stride_h, relu, act)
    assert y.shape[:3] == ref.shape[:3]
    got = y[..., :Cout]
    scale = float(np.abs(ref).max())
    tol = 2e-3 * scale
    if not out_fp32:
        tol += scale * (2.0 ** -8 if act == 1 else 2.0 ** -11)
    err = float(np.abs(got - ref).max())
    assert err <= tol, "%s: max abs err %g > tol %g (scale %g)" % (name, err, tol, scale)
    if y_extra:
        assert np.all(y[..., Cout:] == 0), "kernel wrote outside its channel view"


POOL_CASES = [
    # name, B, H, W, Cin, Cout, act, n_tile   (conv3x3 p1 + ReLU + MaxPool2d(2,2), the vgg / ResNet / TPS pattern)
    ("vgg_c64", 1, 64, 96, 64, 64, 0, 0),
    ("vgg_c128_m256", 2, 96, 128, 64, 128, 0, 0),
    ("vgg_c256", 1, 40, 24, 128, 256, 0, 0),
    ("vgg_c512_n256", 1, 20, 12, 64, 512, 1, 0),
    ("crnn_32x100", 6, 32, 100, 32, 64, 0, 0),
    ("crnn_16x50_odd_w", 9, 16, 50, 128, 128, 0, 0),
    ("odd_hw", 3, 9, 25, 64, 64, 1, 0),
    ("vgg_c64_ragged", 2, 50, 70, 64, 64, 0, 0),
    ("c32_sw64", 2, 16, 36, 32, 32, 0, 0),
    ("cta2_vgg_c128", 4, 160, 128, 128, 128, 0, 0),      # the slice1.10 pattern on the CTA-pair kernels
    ("cta2_crnn_conv1", 125, 16, 50, 128, 128, 0, 0),
    ("halo_c64_large", 2, 160, 128, 64, 64, 0, 0),       # the slice1.3 pattern over more tiles than SMs
    ("halo_c64_odd", 1, 336, 240, 64, 64, 1, 0),
    ("cta2_n256_vgg_c256", 2, 80, 60, 256, 256, 0, 0),   # the slice3.20 pattern (N = 256 tiles on CTA pairs)
    ("cta2_n256_vgg_c512", 2, 40, 30, 512, 512, 1, 0),
]


@pytest.mark.parametrize("case", POOL_CASES, ids=[c[0] for c in POOL_CASES])
@pytest.mark.parametrize("want_full", [True, False])
def test_conv_fused_maxpool(case, want_full):
    """Pooled output == max_pool2d(kernel's own full-resolution output), bit for bit (max commutes with rounding),
    and the full-resolution output is unchanged by the fusion."""
    from lightly_ocr_b200 import bridge
    name, B, H, W, Cin, Cout, act, n_tile = case
    import zlib
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    x = rng.standard_normal((B, H, W, Cin)).astype(np.float32)
    w = (rng.standard_normal((Cout, 3, 3, Cin)) / np.sqrt(9 * Cin)).astype(np.float32)
    bias = rng.standard_normal(Cout).astype(np.float32)
    plain = bridge.test_conv(x, w, bias, None, pad=(1, 1), relu=True, out_fp32=False, act_dtype=act, n_tile=n_tile)
    y, yp = bridge.test_conv_pool(x, w, bias, pad=(1, 1), relu=True, act_dtype=act, n_tile=n_tile, want_full=want_full)
    want = F.max_pool2d(torch.from_numpy(plain).permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).numpy()
    assert yp.shape == want.shape
    assert np.array_equal(yp, want), "%s: %d pooled values differ" % (name, int((yp != want).sum()))
    if want_full:
        assert np.array_equal(y, plain)


@pytest.mark.parametrize("shape", [(1, 16, 16), (2, 33, 47), (1, 160, 120), (3, 20, 100), (2, 176, 128), (1, 330, 250)],
                         ids=str)
@pytest.mark.parametrize("act", [0, 1])
def test_conv_halo_mode_64_to_64(shape, act):
    """3x3 / pad 1 / 64 -> 64 layers take the haloed-patch path (one TMA patch per 16 x 16 tile, the nine taps are
    shared-memory descriptor offsets): same tolerance as every other layer, incl. image borders and ragged edges."""
    from lightly_ocr_b200 import bridge
    B, H, W = shape
    rng = np.random.default_rng(B * 1000 + H * 10 + W + act)
    x = rng.standard_normal((B, H, W, 64)).astype(np.float32)
    w = (rng.standard_normal((64, 3, 3, 64)) / 24.0).astype(np.float32)
    bias = rng.standard_normal(64).astype(np.float32)
    y = bridge.test_conv(x, w, bias, None, pad=(1, 1), relu=True, out_fp32=False, act_dtype=act)
    ref = _ref(x, w, bias, None, (1, 1), (1, 1), 1, True, act)
    scale = float(np.abs(ref).max())
    tol = 2e-3 * scale + scale * (2.0 ** -8 if act == 1 else 2.0 ** -11)
    err = float(np.abs(y - ref).max())
    assert err <= tol, "max abs err %g > tol %g" % (err, tol)


@pytest.mark.parametrize("chan", [(64, 32), (32, 64), (32, 32), (32, 16)], ids=str)
@pytest.mark.parametrize("shape", [(1, 16, 16), (2, 33, 47), (1, 80, 60), (3, 20, 100)], ids=str)
@pytest.mark.parametrize("act", [0, 1])
def test_conv_halo_mode_narrow_layers(shape, chan, act):
    """The haloed-patch form on the decoder-tail shapes: 64 -> 32 channels (128-byte pixels, N = 32) and 32 -> 64 / 32 /
    16 channels (64-byte pixels: the nine taps as descriptor offsets into a 64-byte-swizzled patch)."""
    from lightly_ocr_b200 import bridge
    B, H, W = shape
    cin, cout = chan
    rng = np.random.default_rng(B * 1000 + H * 10 + W + act + 7 * cin + cout)
    x = rng.standard_normal((B, H, W, cin)).astype(np.float32)
    w = (rng.standard_normal((cout, 3, 3, cin)) / np.sqrt(9 * cin)).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32)
    y = bridge.test_conv(x, w, bias, None, pad=(1, 1), relu=True, out_fp32=False, act_dtype=act)
    ref = _ref(x, w, bias, None, (1, 1), (1, 1), 1, True, act)
    scale = float(np.abs(ref).max())
    tol = 2e-3 * scale + scale * (2.0 ** -8 if act == 1 else 2.0 ** -11)
    err = float(np.abs(y - ref).max())
    assert err <= tol, "max abs err %g > tol %g" % (err, tol)


@pytest.mark.parametrize("cin", [64, 128, 256])
@pytest.mark.parametrize("shape", [(2, 160, 128), (1, 336, 240), (4, 96, 112), (1, 330, 250)], ids=str)
@pytest.mark.parametrize("pool", [False, True])
def test_conv_hstream_pairs(shape, cin, pool):
    """The 128 -> 128 3x3 layers at high resolution (slice1.10) run as CTA pairs
    over haloed patches (one per 64-channel chunk) against a streamed weight ring: even and odd tile counts, ragged
    edges within the 5 % waste limit, with and without the fused 2x2 max-pool.  64 / 256 input channels: the generic pair
    form, same shapes."""
    from lightly_ocr_b200 import bridge
    B, H, W = shape
    rng = np.random.default_rng(B * 1000 + H * 10 + W + cin)
    x = rng.standard_normal((B, H, W, cin)).astype(np.float32)
    w = (rng.standard_normal((128, 3, 3, cin)) / np.sqrt(9 * cin)).astype(np.float32)
    bias = rng.standard_normal(128).astype(np.float32)
    ref = _ref(x, w, bias, None, (1, 1), (1, 1), 1, True, 0)
    scale = float(np.abs(ref).max())
    tol = 2e-3 * scale + scale * 2.0 ** -11
    if pool:
        y, yp = bridge.test_conv_pool(x, w, bias, pad=(1, 1), relu=True, act_dtype=0, want_full=True)
        want = F.max_pool2d(torch.from_numpy(y).permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).numpy()
        assert np.array_equal(yp, want)
    else:
        y = bridge.test_conv(x, w, bias, None, pad=(1, 1), relu=True, out_fp32=False, act_dtype=0)
    err = float(np.abs(y - ref).max())
    assert err <= tol, "max abs err %g > tol %g" % (err, tol)


def test_cta_pair_kernels_on_every_eligible_layer():
    """By default only the K >= 1152 layers with N = 128 take the CTA-pair kernels (cta_group::2).  The switch is read
    once per process, so the `cta2_*` cases above are re-run in a child process with LOCR_CONV_CTA2=2, where every
    eligible layer - small K, 1x1, residual, fused pools, odd tile counts, and (LOCR_CONV_CTA2_N256=2) the N = 256 tiles of
    the 256- / 512- / 1024-channel layers - goes through them; LOCR_CONV_HSTREAM=0 there, so the shapes that take the
    haloed-patch pair form by default are covered in the generic pair form as well."""
    import os
    import subprocess
    import sys
    if os.environ.get("LOCR_CONV_CTA2") == "2":
        pytest.skip("already inside the child process")
    env = dict(os.environ, LOCR_CONV_CTA2="2", LOCR_CONV_CTA2_N256="2", LOCR_CONV_HSTREAM="0")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-k", "(cta2 or hstream) and not every_eligible"],
                       env=env, capture_output=True, text=True, timeout=600,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    tail = (r.stdout + r.stderr)[-2000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail



SPLITK_CASES = [
    # name, B, H, W, Cin, Cout, KH, KW, pad, stride_h, residual, act   (16-bit outputs, ReLU; <= 1024 output pixels)
    ("one_crop_c512", 1, 4, 26, 512, 512, 3, 3, (1, 1), 1, True, 0),          # 24 of the recogniser's 29 convs at B = 1
    ("three_crops_c512", 3, 4, 26, 512, 512, 3, 3, (1, 1), 1, True, 0),       # three m-tiles: four slices
    ("eight_crops_c512_bf16", 8, 4, 26, 512, 512, 3, 3, (1, 1), 1, False, 1),  # 7 m-tiles: two slices
    ("one_crop_c256", 1, 8, 25, 256, 256, 3, 3, (1, 1), 1, True, 0),          # layer2: four chunks -> four slices
    ("one_crop_c128", 1, 16, 50, 128, 128, 3, 3, (1, 1), 1, True, 0),         # layer1: two chunks -> two slices
    ("one_crop_c256_c512", 1, 4, 26, 256, 512, 3, 3, (1, 1), 1, False, 0),    # layer3.0.conv1
    ("conv4_1_k2_s21", 1, 4, 26, 512, 512, 2, 2, (0, 1), 2, False, 0),        # k2 s(2,1) p(0,1)
    ("conv4_2_k2", 2, 2, 27, 512, 512, 2, 2, (0, 0), 1, False, 0),            # k2 s1 p0
    ("downsample_1x1_c512", 1, 4, 26, 512, 512, 1, 1, (0, 0), 1, False, 0),   # 8 k-blocks
    ("views_c64_of_c192", 2, 10, 20, 128, 64, 3, 3, (1, 1), 1, True, 0),      # channel-slice input / output views
]


@pytest.mark.parametrize("case", SPLITK_CASES, ids=[c[0] for c in SPLITK_CASES])
def test_conv_split_k(case, monkeypatch):
    """Split-K form of the deep layers on single crops / a handful of them (conv_tc.cuh splitk_ws): the K range runs as
    up to 8 slices x n-tiles CTAs that store fp32 partial sums, a second kernel adds slices + bias + residual, applies
    the ReLU and rounds.  Against torch's fp32 conv2d on the same rounded operands (tolerance of test_conv_parity), and
    against the unsplit kernel: the two only differ in the order of fp32 additions, i.e. by at most one 16-bit
    rounding step of a value that sat on a rounding boundary."""
    from lightly_ocr_b200 import bridge
    name, B, H, W, Cin, Cout, KH, KW, pad, stride_h, use_res, act = case
    import zlib
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    x_extra, y_extra = (64, 32) if name.startswith("views") else (0, 0)
    x = rng.standard_normal((B, H, W, Cin + x_extra)).astype(np.float32)
    w = (rng.standard_normal((Cout, KH, KW, Cin)) / np.sqrt(KH * KW * Cin)).astype(np.float32)
    bias = rng.standard_normal(Cout).astype(np.float32)
    OH = (H + 2 * pad[0] - (KH - 1) - 1) // stride_h + 1
    OW = W + 2 * pad[1] - (KW - 1)
    residual = rng.standard_normal((B, OH, OW, Cout)).astype(np.float32) if use_res else None
    L = bridge.lib()
    L.locr_test_splitk_calls.restype = __import__("ctypes").c_int64
    kw = dict(pad=pad, stride_h=stride_h, relu=True, out_fp32=False, act_dtype=act, y_pitch=Cout + y_extra)
    monkeypatch.setenv("LOCR_TEST_SPLITK", "1")
    n0 = L.locr_test_splitk_calls()
    y = bridge.test_conv(x, w, bias, residual, **kw)
    assert L.locr_test_splitk_calls() == n0 + 1, "%s did not take the split-K form" % name
    monkeypatch.setenv("LOCR_TEST_SPLITK", "0")
    plain = bridge.test_conv(x, w, bias, residual, **kw)
    assert L.locr_test_splitk_calls() == n0 + 1
    ref = _ref(x[..., :Cin], w, bias, residual, (1, 1), pad, stride_h, True, act)
    scale = float(np.abs(ref).max())
    ulp = scale * (2.0 ** -8 if act == 1 else 2.0 ** -11)
    err = float(np.abs(y[..., :Cout] - ref).max())
    assert err <= 2e-3 * scale + ulp, "%s: max abs err %g (scale %g)" % (name, err, scale)
    diff = np.abs(y - plain)
    assert float(diff.max()) <= 2 * ulp, "%s: split and unsplit results differ by %g" % (name, float(diff.max()))
    assert float((diff > 0).mean()) < 0.02
    if y_extra:
        assert np.all(y[..., Cout:] == 0), "kernel wrote outside its channel view"
