"""GPU parity tests, operator level (several cases exceed 148 tiles so that every persistent CTA loops over more
than one tile — barrier phase wrap-around, both TMEM accumulator stages): each CUDA kernel through the C ABI vs the oracle on the same seeded
inputs.  Integer outputs (labels) must be bit exact; bf16 outputs within one bf16 rounding of the fp32 result."""
import os

import numpy as np
import pytest

from dlv3p_b200 import ffi
from oracle import head_ref as R
from tests.common import rel_err

pytestmark = pytest.mark.gpu

BF16_EPS = 2.0 ** -8     # half ulp of bf16 relative to the value (8 bit mantissa incl. hidden bit)


def _assert_bf16_close(got_bits, ref_f32, extra_abs=0.0, what=''):
    got = ffi.bf16_bits_to_f32(got_bits)
    tol = np.abs(ref_f32) * (2 * BF16_EPS) + extra_abs
    bad = np.abs(got - ref_f32) > tol
    assert not bad.any(), '%s: %d / %d elements out of tolerance, max abs err %.4g (rel-to-max %.3g)' % (
        what, bad.sum(), bad.size, np.abs(got - ref_f32).max(), rel_err(got, ref_f32))


@pytest.mark.parametrize('M,K,N', [(128, 64, 256), (256, 2048, 256), (1000, 304, 256), (4096, 256, 48), (777, 320, 256),
                                    (512, 256, 24), (300, 1024, 256), (130, 160, 256), (64, 96, 16), (20000, 256, 256),
                                    (128 * 148 * 3 + 5, 64, 256), (128 * 148 * 5 + 77, 128, 48), (128 * 148 * 2 + 1, 320, 256)])
def test_pointwise_gemm(gpu, M, K, N):
    """tcgen05 1x1 conv + BN + ReLU (layers.py:14-21, :141-143) incl. ragged M, K % 64 != 0, narrow N."""
    rng = np.random.default_rng(M + K + N)
    a = R.bf16_round(rng.standard_normal((M, K)).astype(np.float32))
    w = rng.standard_normal((K, N)).astype(np.float32) * np.float32(np.sqrt(2.0 / K))
    s = rng.uniform(0.5, 1.5, N).astype(np.float32)
    t = rng.standard_normal(N).astype(np.float32) * 0.1
    got = ffi.op_pointwise(R.to_bf16_bits(a), w, s, t, relu=True)
    ref = np.maximum((a.astype(np.float64) @ R.bf16_round(w).astype(np.float64)) * s + t, 0).astype(np.float32)
    _assert_bf16_close(got, ref, extra_abs=2e-3, what='pointwise M=%d K=%d N=%d' % (M, K, N))
    # no-ReLU / no-BN variant: pure GEMM
    got2 = ffi.op_pointwise(R.to_bf16_bits(a), w, None, None, relu=False)
    ref2 = (a.astype(np.float64) @ R.bf16_round(w).astype(np.float64)).astype(np.float32)
    _assert_bf16_close(got2, ref2, extra_abs=2e-3, what='gemm M=%d K=%d N=%d' % (M, K, N))


@pytest.mark.parametrize('B,H,W,C,rate', [(2, 32, 32, 64, 6), (1, 32, 32, 128, 18), (1, 20, 17, 24, 1), (1, 40, 33, 8, 12),
                                           (2, 16, 16, 304, 1), (1, 64, 48, 16, 36)])
def test_depthwise(gpu, B, H, W, C, rate):
    """dilated depthwise 3x3 'same' + BN + ReLU (layers.py:100-104)."""
    rng = np.random.default_rng(B * 1000 + H + C + rate)
    x = R.bf16_round(rng.standard_normal((B, H, W, C)).astype(np.float32))
    k = rng.standard_normal((3, 3, C, 1)).astype(np.float32) * 0.3
    s = rng.uniform(0.5, 1.5, C).astype(np.float32)
    t = rng.standard_normal(C).astype(np.float32) * 0.1
    got = ffi.op_depthwise(R.to_bf16_bits(x), k[..., 0], rate, s, t, relu=True)
    ref = np.maximum(R.depthwise3x3(x, k, rate) * s + t, 0)
    _assert_bf16_close(got, ref, extra_abs=1e-5, what='depthwise')


@pytest.mark.parametrize('B,H,W,C', [(1, 8, 16, 64), (2, 16, 32, 256), (1, 24, 40, 304), (1, 13, 21, 128), (3, 9, 7, 192),
                                      (1, 64, 64, 304), (2, 32, 32, 320), (1, 5, 3, 8),
                                      (4, 64, 128, 304), (8, 128, 128, 256), (6, 100, 76, 304), (5, 72, 88, 64)])
def test_fused_sepconv(gpu, B, H, W, C):
    """SepConv_BN(depth_activation=True) fused: depthwise stencil as the on-chip A operand of the tcgen05 GEMM
    (layers.py:74-111; decoder_conv0 C=304, decoder_conv1 C=256)."""
    rng = np.random.default_rng(B + H * 7 + W * 13 + C)
    x = R.bf16_round(rng.standard_normal((B, H, W, C)).astype(np.float32))
    dk = rng.standard_normal((3, 3, C, 1)).astype(np.float32) * 0.3
    ds = rng.uniform(0.5, 1.5, C).astype(np.float32)
    dt = rng.standard_normal(C).astype(np.float32) * 0.1
    pk = rng.standard_normal((C, 256)).astype(np.float32) * np.float32(np.sqrt(2.0 / C))
    ps = rng.uniform(0.5, 1.5, 256).astype(np.float32)
    pt = rng.standard_normal(256).astype(np.float32) * 0.1
    got = ffi.op_sepconv(R.to_bf16_bits(x), dk[..., 0], ds, dt, pk, ps, pt)
    mid = R.bf16_round(np.maximum(R.depthwise3x3(x, dk, 1) * ds + dt, 0))
    ref = np.maximum((mid.reshape(-1, C).astype(np.float64) @ R.bf16_round(pk).astype(np.float64)) * ps + pt, 0)
    ref = ref.reshape(B, H, W, 256).astype(np.float32)
    # the intermediate is rounded to bf16 on both sides; a 1-ulp flip there moves the output by ~|w|*ulp
    got_f = ffi.bf16_bits_to_f32(got)
    assert rel_err(got_f, ref) < 6e-3, 'fused sepconv rel err %.3g' % rel_err(got_f, ref)
    frac_bad = (np.abs(got_f - ref) > np.abs(ref) * (4 * BF16_EPS) + 2e-2).mean()
    assert frac_bad < 1e-3, 'fused sepconv: %.4f of elements off' % frac_bad


@pytest.mark.parametrize('B,hi,wi,C,ho,wo', [(2, 32, 32, 256, 128, 128), (1, 7, 5, 16, 25, 19), (1, 1, 1, 8, 9, 4), (1, 16, 16, 24, 7, 9),
                                              (1, 33, 33, 8, 129, 129)])
def test_resize_bilinear(gpu, B, hi, wi, C, ho, wo):
    """tf.image.resize bilinear, half-pixel centres (layers.py:48-50): same op order as the oracle -> bit exact."""
    rng = np.random.default_rng(hi * wi + C)
    x = R.bf16_round(rng.standard_normal((B, hi, wi, C)).astype(np.float32))
    got = ffi.op_resize_bilinear(R.to_bf16_bits(x), ho, wo)
    ref = R.to_bf16_bits(R.resize_bilinear(x, (ho, wo)))
    assert np.array_equal(got, ref), 'resize mismatch on %d elements' % (got != ref).sum()


@pytest.mark.parametrize('B,NC,hi,wi,ho,wo', [(2, 21, 128, 128, 512, 512), (1, 19, 25, 19, 100, 76), (1, 21, 32, 32, 512, 512),
                                               (1, 5, 9, 7, 36, 28), (1, 150, 16, 16, 64, 64), (1, 3, 10, 10, 33, 47), (1, 21, 1, 1, 4, 4),
                                               (2, 21, 16, 24, 128, 192), (1, 19, 8, 8, 256, 256), (1, 4, 5, 3, 60, 36)])
def test_resize_argmax_bit_exact(gpu, B, NC, hi, wi, ho, wo):
    """pred_resize + np.argmax (model.py:76, deeplab.py:99): integer output, bit exact incl. first-max ties."""
    rng = np.random.default_rng(NC * hi + wo)
    logits = rng.standard_normal((B, hi, wi, NC)).astype(np.float32)
    # plant exact ties: duplicate a class plane into a later class
    if NC >= 4:
        logits[..., NC - 1] = logits[..., 1]
        logits[..., 2] = np.maximum(logits[..., 2], logits[..., 1])
    ref = R.argmax_labels(R.resize_bilinear(logits, (ho, wo))).astype(np.uint8)
    got = ffi.op_resize_argmax(np.ascontiguousarray(np.transpose(logits, (0, 3, 1, 2))), ho, wo)
    assert np.array_equal(got, ref), 'labels differ on %d of %d pixels' % ((got != ref).sum(), ref.size)


def test_ops_reject_bad_arguments(gpu):
    import dlv3p_b200
    with pytest.raises(dlv3p_b200.Dlv3pError):
        ffi.op_pointwise(np.zeros((128, 60), np.uint16), np.zeros((60, 256), np.float32))      # K % 8
    with pytest.raises(dlv3p_b200.Dlv3pError) as e:
        ffi.op_sepconv(np.zeros((1, 8, 8, 64), np.uint16), np.zeros((3, 3, 64)), np.ones(64), np.zeros(64),
                       np.zeros((64, 256)), np.ones(256), np.zeros(256), rate=2)                # fused kernel: rate 1 only
    assert e.value.status == -3


def test_fused_sepconv_is_deterministic_and_batch_independent(gpu):
    """Many tiles per persistent CTA, identical images in every batch slot: any barrier-phase race between the
    stencil warps shows up as slot-to-slot or run-to-run differences (integer comparison of the bf16 bit patterns)."""
    rng = np.random.default_rng(2)
    for C in (256, 304):
        x1 = R.to_bf16_bits(rng.standard_normal((2, 128, 128, C)).astype(np.float32))
        x = np.tile(x1, (12, 1, 1, 1))
        dk = rng.standard_normal((3, 3, C)).astype(np.float32) * 0.3
        pk = rng.standard_normal((C, 256)).astype(np.float32) * np.float32(np.sqrt(2.0 / C))
        args = (dk, np.ones(C, np.float32), np.zeros(C, np.float32), pk, np.ones(256, np.float32), np.zeros(256, np.float32))
        a = ffi.op_sepconv(x, *args)
        b = ffi.op_sepconv(x, *args)
        assert np.array_equal(a, b), 'C=%d: run-to-run difference on %d elements' % (C, (a != b).sum())
        for rep in range(1, 12):
            assert np.array_equal(a[:2], a[2 * rep:2 * rep + 2]), 'C=%d: batch slot %d differs from slot 0' % (C, rep)


@pytest.mark.parametrize('nc,n', [(21, 512 * 512), (2, 1000), (150, 33 * 47), (256, 4096), (64, 17)])
def test_confusion_matrix_bit_exact(gpu, nc, n):
    """eval.py:368-373 generate_matrix (+ the running sum over images): integer work, bit exact, ignore label 255."""
    rng = np.random.default_rng(nc + n)
    gt = rng.integers(0, nc, size=n).astype(np.uint8)
    pr = np.where(rng.random(n) < 0.8, gt, rng.integers(0, nc, size=n)).astype(np.uint8)
    if nc < 255:
        gt[rng.random(n) < 0.05] = 255
    ref = R.generate_matrix(gt, pr, nc)
    assert np.array_equal(ffi.op_confusion_matrix(pr, gt, nc), ref)
    assert np.array_equal(ffi.op_confusion_matrix(pr, gt, nc, repeat=3), 3 * ref)


def test_confusion_matrix_equals_the_reference_output(gpu):
    """Against the output of the reference's own generate_matrix (tests/golden/ref_pins.npz, make_ref_pins.py)."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_pins.npz'))
    for k in range(4):
        got = ffi.op_confusion_matrix(z['pred_%d' % k], z['gt_%d' % k], int(z['nc_%d' % k]))
        assert np.array_equal(got, z['confusion_%d' % k])


# ---------------------------------------------------------------------------------------------------- N4: image pre / post-processing
@pytest.mark.gpu
def test_image_pre_post_processing_matches_the_reference_outputs(gpu):
    """normalize_image / denormalize_image / mask_resize on the device, bit exact against outputs of the reference's OWN functions
    (tests/golden/ref_pins.npz, generated by executing common/data_utils.py:403-477 — see tests/golden/make_ref_pins.py)."""
    from dlv3p_b200 import ffi
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_pins.npz'))
    assert np.array_equal(ffi.op_normalize_image(z['image_u8']), z['normalize'])
    assert np.array_equal(ffi.op_denormalize_image(z['normalize']), z['denormalize'])
    assert np.array_equal(ffi.op_mask_resize(z['mask_in'], (27, 36)), z['mask_resize_27x36'])
    # every uint8 value; bf16 output = round-to-nearest-even of the fp32 result (the head's input dtype)
    allv = np.arange(256, dtype=np.uint8)
    ref = allv.astype(np.float32) / 127.5 - 1
    assert np.array_equal(ffi.op_normalize_image(allv), ref)
    assert np.array_equal(ffi.op_normalize_image(allv, out_bf16=True), ffi.f32_to_bf16_bits(ref))
    assert np.array_equal(ffi.op_denormalize_image(ref), (ref * 127.5 + 127.5).astype(np.uint8))


@pytest.mark.gpu
@pytest.mark.parametrize('hi,wi,ho,wo', [(512, 512, 375, 500), (512, 512, 1024, 2048), (128, 96, 13, 7), (5, 7, 500, 333), (1, 1, 9, 9), (64, 64, 64, 64)])
def test_mask_resize_is_cv2_inter_nearest(gpu, hi, wi, ho, wo):
    """The reference's mask_resize IS cv2.resize(INTER_NEAREST) (common/data_utils.py:476): compare with cv2 itself, batched."""
    import cv2
    from dlv3p_b200 import ffi
    rng = np.random.default_rng(hi * 7 + wo)
    masks = rng.integers(0, 21, size=(3, hi, wi)).astype(np.uint8)
    got = ffi.op_mask_resize(masks, (wo, ho))
    for b in range(3):
        assert np.array_equal(got[b], cv2.resize(masks[b], (wo, ho), interpolation=cv2.INTER_NEAREST))


@pytest.mark.gpu
@pytest.mark.parametrize('hw_in,hw_out,B', [((64, 48), (32, 32), 2), ((37, 51), (64, 80), 1), ((375, 500), (512, 512), 2), ((512, 512), (512, 300), 1),
                                            ((96, 96), (96, 96), 1), ((9, 7), (40, 3), 3), ((300, 200), (31, 29), 1), ((33, 65), (16, 65), 2),
                                            ((1024, 2048), (512, 512), 1)])
def test_resize_in_is_pil_bicubic(gpu, hw_in, hw_out, B):
    """The resize of preprocess_image (common/data_utils.py:449, PIL Image.resize(..., Image.BICUBIC)) on the device: bit exact against
    the oracle's restatement of Pillow's resampler (itself pinned against Pillow) and against Pillow directly, batched."""
    from PIL import Image
    from dlv3p_b200 import ffi
    rng = np.random.default_rng(hw_in[0] * 131 + hw_out[1])
    img = rng.integers(0, 256, (B,) + hw_in + (3,)).astype(np.uint8)
    img[:, : hw_in[0] // 2] = (img[:, : hw_in[0] // 2] // 128) * 255     # saturated blocks: the cubic's overshoot hits clip8
    got = ffi.op_resize_bicubic(img, hw_out)
    assert got.shape == (B,) + hw_out + (3,)
    for b in range(B):
        assert np.array_equal(got[b], R.pil_bicubic_resize(img[b], hw_out))
        assert np.array_equal(got[b], np.asarray(Image.fromarray(img[b]).resize((hw_out[1], hw_out[0]), Image.BICUBIC)))
    # single channel (a grey image) goes through the same passes
    g = ffi.op_resize_bicubic(img[0, :, :, :1], hw_out)
    assert np.array_equal(g[..., 0], np.asarray(Image.fromarray(img[0, :, :, 0]).resize((hw_out[1], hw_out[0]), Image.BICUBIC)))


@pytest.mark.gpu
def test_present_classes_of_the_native_postprocess(gpu):
    """class_indexes (inference/MNN/deeplabSegment.cpp:171-172): classes != 0 in order of first appearance, per image."""
    from dlv3p_b200 import ffi
    rng = np.random.default_rng(11)
    masks = rng.integers(0, 21, (3, 375, 500)).astype(np.uint8)
    masks[0, :200] = 0
    masks[0, 200, 17] = 20
    masks[1] = 0
    masks[2, 0, 0] = 255
    got = ffi.op_present_classes(masks)
    assert got == [R.present_classes(m) for m in masks]
    assert got[0][0] == 20 and got[1] == [] and got[2][0] == 255
    big = np.zeros((1, 1024, 2048), np.uint8)
    big[0, -1, -1] = 3
    big[0, 5, 5] = 9
    assert ffi.op_present_classes(big) == [[9, 3]]
