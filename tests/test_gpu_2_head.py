"""GPU parity tests, whole hot path through the C ABI vs the oracle and the committed golden vectors.

Tolerances (BASELINE.json north_star): logits within 1e-2 relative (max|d|/max|ref|) of the fp32 oracle in bf16;
argmax labels identical on >= 99.9 % of pixels against the oracle run with the same bf16 rounding points
(against the pure-fp32 oracle random-weight logits have thin margins -> reported, looser bound)."""
import numpy as np
import pytest

from dlv3p_b200 import ffi
from oracle import head_ref as R
from tests.common import GOLDEN, HEAD_CASES, label_agreement, load_case, make_head, planar_to_nhwc, rel_err

pytestmark = pytest.mark.gpu

LOGIT_TOL = 1e-2
LABEL_AGREE = 0.999


def check_labels(labels, o16, what):
    """>= 99.9 % identical labels vs the bf16-sim oracle wherever the oracle's own top-1 margin exceeds the logits
    tolerance; every mismatch must sit on a near-tie (margin < tolerance); overall agreement reported and bounded."""
    ref_labels = o16['labels'].numpy() if hasattr(o16['labels'], 'numpy') else o16['labels']
    full = o16['logits_full'].numpy() if hasattr(o16['logits_full'], 'numpy') else o16['logits_full']
    overall, decided, worst = label_agreement(labels, ref_labels, full, LOGIT_TOL)
    print('%s: label agreement overall %.5f, on decided pixels %.5f, worst mismatch margin %.4f' % (what, overall, decided, worst))
    assert decided >= LABEL_AGREE, '%s: agreement on decided pixels %.5f' % (what, decided)
    assert worst < LOGIT_TOL, '%s: a mismatch at a pixel with oracle margin %.4f' % (what, worst)
    assert overall >= 0.995, '%s: overall agreement %.5f' % (what, overall)
    return overall


@pytest.mark.parametrize('name', HEAD_CASES)
def test_head_matches_golden(gpu, name):
    cfg, W, feat, skip, z = load_case(name)
    hd = make_head(cfg, W, out_mode=ffi.OUT_LABELS_U8)
    labels = hd(feat, skip)
    logits = planar_to_nhwc(hd.tap('logits'))
    assert rel_err(logits, z['logits_fp32']) < LOGIT_TOL, 'logits vs fp32 oracle: %.3g' % rel_err(logits, z['logits_fp32'])
    assert rel_err(logits, z['logits_bf16']) < 8e-3, 'logits vs bf16-sim oracle: %.3g' % rel_err(logits, z['logits_bf16'])
    agree = (labels == z['labels_bf16']).mean()
    assert agree >= LABEL_AGREE, 'label agreement vs bf16-sim oracle %.5f' % agree
    assert (labels == z['labels_fp32']).mean() >= 0.98
    hd.close()


@pytest.mark.parametrize('name', ['head_small_full', 'head_odd_size', 'head_small_lite_dec'])
def test_intermediate_taps(gpu, name):
    """Every block boundary against the oracle taps keyed by reference layer names."""
    cfg, W, feat, skip, _ = load_case(name)
    ref = R.head_forward(feat, skip, W, cfg, 'bf16')
    hd = make_head(cfg, W)
    hd(feat, skip)
    assert rel_err(hd.tap('image_pooling'), ref['image_pooling']) < 1e-2
    assert rel_err(hd.tap('aspp_out'), ref['aspp_out']) < 8e-3
    assert rel_err(hd.tap('decoder_in'), ref['decoder_in']) < 8e-3
    assert rel_err(hd.tap('decoder_conv0'), ref['decoder_conv0']) < 8e-3
    assert rel_err(hd.tap('decoder_out'), ref['decoder_out']) < 8e-3
    hd.close()


def test_fused_and_unfused_decoder_agree(gpu):
    cfg, W, feat, skip, _ = load_case('head_small_full')
    a = make_head(cfg, W, out_mode=ffi.OUT_LOGITS_LOWRES)
    b = make_head(cfg, W, out_mode=ffi.OUT_LOGITS_LOWRES, flags=ffi.FLAG_UNFUSED_DECODER)
    la, lb = a(feat, skip), b(feat, skip)
    assert rel_err(la, lb) < 4e-3
    a.close(); b.close()


def test_output_modes_consistent(gpu):
    cfg, W, feat, skip, z = load_case('head_small_full')
    labels = make_head(cfg, W, out_mode=ffi.OUT_LABELS_U8)(feat, skip)
    low = make_head(cfg, W, out_mode=ffi.OUT_LOGITS_LOWRES)(feat, skip)
    full = make_head(cfg, W, out_mode=ffi.OUT_LOGITS_FULL)(feat, skip)
    prob = make_head(cfg, W, out_mode=ffi.OUT_SOFTMAX)(feat, skip)
    # pred_resize of the low-res logits, bit exact against the oracle's resize of the SAME logits
    assert np.array_equal(full, R.resize_bilinear(planar_to_nhwc(low), (cfg.H, cfg.W)))
    assert np.array_equal(labels, np.argmax(full, -1).astype(np.uint8))
    assert np.abs(prob - R.softmax(full)).max() < 1e-6 and np.abs(prob.sum(-1) - 1).max() < 1e-5
    # reference path: argmax of the softmax output (deeplab.py:99); may differ from logits-argmax only on fp32 softmax ties
    assert (np.argmax(prob, -1) == labels).mean() > 0.9999


def test_fp32_inputs_are_cast_on_device(gpu):
    cfg, W, feat, skip, z = load_case('head_small_full')
    a = make_head(cfg, W, in_dtype=ffi.DTYPE_BF16)(feat, skip)
    b = make_head(cfg, W, in_dtype=ffi.DTYPE_FP32)(feat, skip)
    assert np.array_equal(a, b)


def test_fp16_inputs_are_cast_on_device(gpu):
    """The reference's --mixed_precision hands fp16 activations over: fp16 -> bf16 on the device, same labels as feeding
    the fp16-rounded values through the bf16 path."""
    cfg, W, feat, skip, z = load_case('head_small_full')
    f16, s16 = feat.astype(np.float16), skip.astype(np.float16)
    a = make_head(cfg, W, in_dtype=ffi.DTYPE_BF16)(f16.astype(np.float32), s16.astype(np.float32))
    b = make_head(cfg, W, in_dtype=ffi.DTYPE_FP16)(f16, s16)
    assert np.array_equal(a, b)


def test_forward_host_equals_device_forward_and_is_deterministic(gpu):
    cfg, W, feat, skip, _ = load_case('head_small_os8')
    hd = make_head(cfg, W)
    a = hd(feat, skip)
    b = hd.predict_host(feat, skip)
    c = hd(feat, skip)
    assert np.array_equal(a, b) and np.array_equal(a, c)
    assert hd.ctx.launch_count()[0] >= 9
    hd.close()


def test_forward_host_pipeline_equals_device_forward(gpu):
    """B >= 8: dlv3p_forward_host runs four sub-batches on two child contexts (H2D of one overlaps the kernels of the
    previous); images are independent, so the labels must equal one full-batch device forward bit for bit."""
    cfg = R.HeadConfig(B=8, H=128, W=128, OS=16, Cin=256, Cskip=64, NC=21)
    W = R.make_weights(cfg, 11)
    feat, skip = R.make_inputs(cfg, 12)
    hd = make_head(cfg, W)
    a = hd(feat, skip)
    b = hd.predict_host(feat, skip)
    c = hd.predict_host(feat, skip)
    assert np.array_equal(a, b) and np.array_equal(b, c)
    assert hd.ctx.launch_count()[0] >= 4 * 9       # four sub-batch forwards
    hd.close()


def test_block_level_layers(gpu):
    """ASPP_block / Decoder_block as separate drop-in objects (layers.py:114, :199) chained by the caller."""
    import dlv3p_b200
    cfg, W, feat, skip, _ = load_case('head_small_full')
    ref = R.head_forward(feat, skip, W, cfg, 'bf16')
    aspp = dlv3p_b200.ASPPBlock(feat.shape, cfg.OS)
    aspp.set_weights(W)
    y = aspp(feat)
    assert rel_err(y, ref['aspp_out']) < 8e-3
    dec = dlv3p_b200.DecoderBlock(y.shape, skip.shape)
    dec.set_weights(W)
    d = dec(y, skip)
    assert rel_err(d, ref['decoder_out']) < 8e-3
    with pytest.raises(ValueError):
        dlv3p_b200.ASPPBlock(feat.shape, 4)       # ValueError('invalid output stride', OS), layers.py:126


def test_cfg1_mobilenetv2_os16_512(gpu):
    """BASELINE configs[0]: DeepLabV3+ MobileNetV2 OS16 512x512, 21 classes, batch 1 (the reference's CPU case)."""
    cfg = R.HeadConfig(B=1, H=512, W=512, OS=16, Cin=320, Cskip=24, NC=21)
    W = R.make_weights(cfg, 1234)
    feat, skip = R.make_inputs(cfg, 1235, relu_feat=False)
    hd = make_head(cfg, W)
    labels = hd(feat, skip)
    logits = planar_to_nhwc(hd.tap('logits'))
    o32 = R.head_forward_torch(feat, skip, W, cfg, 'fp32')
    o16 = R.head_forward_torch(feat, skip, W, cfg, 'bf16')
    assert rel_err(logits, o32['logits'].numpy()) < LOGIT_TOL
    check_labels(labels, o16, 'cfg1')
    hd.close()


@pytest.mark.parametrize('kw', [
    dict(B=1, H=136, W=200, OS=16, Cin=96, Cskip=16, NC=150),     # odd tile count (phantom tile of the CTA pair), 150 classes (ADE20K), MobileNetV3-small channels
    dict(B=3, H=96, W=96, OS=8, Cin=64, Cskip=24, NC=2),         # OS8 rates 12/24/36 on a 12x12 map (rates == map size), binary labels
    dict(B=1, H=64, W=64, OS=32, Cin=704, Cskip=128, NC=21),     # OS32 rates 3/6/9 on a 2x2 map, PeleeNet channels
])
def test_edge_geometries(gpu, kw):
    """Geometries at the edges of the kernels' assumptions: partial / phantom tiles, rates >= map size, many or few
    classes, channel counts that are not multiples of 64."""
    cfg = R.HeadConfig(**kw)
    W = R.make_weights(cfg, 31)
    feat, skip = R.make_inputs(cfg, 32, relu_feat=False)
    hd = make_head(cfg, W)
    labels = hd(feat, skip)
    logits = planar_to_nhwc(hd.tap('logits'))
    o32 = R.head_forward_torch(feat, skip, W, cfg, 'fp32')
    o16 = R.head_forward_torch(feat, skip, W, cfg, 'bf16')
    assert rel_err(logits, o32['logits'].numpy()) < LOGIT_TOL
    check_labels(labels, o16, 'edge %s' % (kw,))
    hd.close()


def test_cfg2_xception_os16_512_batch2(gpu):
    """BASELINE configs[1] shapes (Xception OS16 512x512 VOC) at a batch the oracle finishes in seconds."""
    cfg = R.HeadConfig(B=2, H=512, W=512, OS=16, Cin=2048, Cskip=256, NC=21)
    W = R.make_weights(cfg, 1234)
    feat, skip = R.make_inputs(cfg, 1236)
    hd = make_head(cfg, W)
    labels = hd(feat, skip)
    logits = planar_to_nhwc(hd.tap('logits'))
    o32 = R.head_forward_torch(feat, skip, W, cfg, 'fp32')
    o16 = R.head_forward_torch(feat, skip, W, cfg, 'bf16')
    e32 = rel_err(logits, o32['logits'].numpy())
    assert e32 < LOGIT_TOL, 'cfg2 logits vs fp32 oracle %.3g' % e32
    check_labels(labels, o16, 'cfg2')
    hd.close()


def test_cfg2_full_batch_properties(gpu):
    """Full BASELINE size (B=32): size-independent properties instead of an oracle run —
    (1) batch independence: image i of the batch == the same image run alone (inference shards by image),
    (2) run-to-run determinism, (3) label range."""
    cfg = R.HeadConfig(B=32, H=512, W=512, OS=16, Cin=2048, Cskip=256, NC=21)
    W = R.make_weights(cfg, 1234)
    rng = np.random.default_rng(7)
    feat1 = np.maximum(rng.standard_normal((4, 32, 32, 2048), dtype=np.float32), 0)
    skip1 = rng.standard_normal((4, 128, 128, 256), dtype=np.float32)
    feat = np.tile(feat1, (8, 1, 1, 1))
    skip = np.tile(skip1, (8, 1, 1, 1))
    hd = make_head(cfg, W)
    a = hd(feat, skip)
    b = hd(feat, skip)
    assert np.array_equal(a, b)
    assert a.max() < 21
    for rep in range(1, 8):
        assert np.array_equal(a[:4], a[4 * rep:4 * rep + 4]), 'batch slot %d differs from slot 0 on identical inputs' % rep
    cfg1 = R.HeadConfig(B=4, H=512, W=512, OS=16, Cin=2048, Cskip=256, NC=21)
    hd1 = make_head(cfg1, W)
    assert np.array_equal(hd1(feat1, skip1), a[:4])
    hd.close(); hd1.close()


def test_cfg3_cityscapes_os8_shapes(gpu):
    """BASELINE configs[2] (Xception OS8, 19 classes, rates 12/24/36) at reduced spatial size for the oracle,
    non-square map."""
    cfg = R.HeadConfig(B=1, H=256, W=512, OS=8, Cin=2048, Cskip=256, NC=19)
    W = R.make_weights(cfg, 77)
    feat, skip = R.make_inputs(cfg, 78)
    hd = make_head(cfg, W)
    labels = hd(feat, skip)
    o16 = R.head_forward_torch(feat, skip, W, cfg, 'bf16')
    o32 = R.head_forward_torch(feat, skip, W, cfg, 'fp32')
    logits = planar_to_nhwc(hd.tap('logits'))
    assert rel_err(logits, o32['logits'].numpy()) < LOGIT_TOL
    assert rel_err(logits, o16['logits'].numpy()) < 8e-3        # 1-ulp bf16 flips of intermediates on either side
    check_labels(labels, o16, 'cfg3')
    hd.close()


def test_cfg4_mobilenetv3_lite(gpu):
    """BASELINE configs[3]: 4a = reference mobilenetv3large_lite (ASPP Lite, no decoder, x16 resize);
    4b = ASPP Lite + Decoder (BASELINE wording)."""
    for decoder in (False, True):
        cfg = R.HeadConfig(B=2, H=512, W=512, OS=16, Cin=160, Cskip=24, NC=21, lite=True, decoder=decoder)
        W = R.make_weights(cfg, 5)
        feat, skip = R.make_inputs(cfg, 6, relu_feat=False)
        hd = make_head(cfg, W)
        labels = hd(feat, skip)
        o16 = R.head_forward_torch(feat, skip, W, cfg, 'bf16')
        o32 = R.head_forward_torch(feat, skip, W, cfg, 'fp32')
        logits = planar_to_nhwc(hd.tap('logits'))
        assert rel_err(logits, o32['logits'].numpy()) < LOGIT_TOL
        assert rel_err(logits, o16['logits'].numpy()) < 8e-3
        check_labels(labels, o16, 'cfg4%s' % ('b' if decoder else 'a'))
        hd.close()


@pytest.mark.parametrize('stem', ['2007_000039', '2007_000346'])
def test_example_images_identical_miou(gpu, stem):
    """north_star: identical mIoU to 3 decimals on the example images (example/2007_000039, 2007_000346), with seeded
    random weights and a stand-in backbone — the value is meaningless, the EQUALITY oracle == CUDA is the test."""
    import os
    z = np.load(os.path.join(GOLDEN, 'example_%s.npz' % stem))
    img = R.normalize_image(z['image'])[None]                       # data_utils.py:403-417
    cfg = R.HeadConfig(B=1, H=256, W=256, OS=16, Cin=320, Cskip=24, NC=21)
    feat, skip = R.standin_backbone(img, cfg)
    W = R.make_weights(cfg, 2024)
    hd = make_head(cfg, W)
    labels = hd(feat, skip)[0]
    o16 = R.head_forward_torch(feat, skip, W, cfg, 'bf16')
    ref = o16['labels'].numpy()[0]
    check_labels(labels[None], o16, 'example ' + stem)
    gt = z['label']
    assert round(R.mIOU(gt, labels), 3) == round(R.mIOU(gt, ref), 3)
    cm_a, cm_b = R.generate_matrix(gt, labels.astype(np.int64), 21), R.generate_matrix(gt, ref.astype(np.int64), 21)
    assert round(R.dataset_mIOU(cm_a), 3) == round(R.dataset_mIOU(cm_b), 3)
    hd.close()
