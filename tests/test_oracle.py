"""CPU tests that pin the oracle: (1) each restated op against an independent implementation
(torch.nn.functional), (2) the committed golden vectors, (3) algebraic properties."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import head_ref as R
from tests.common import HEAD_CASES, load_case, rel_err


def _nchw(x):
    return torch.from_numpy(np.ascontiguousarray(x)).permute(0, 3, 1, 2)


@pytest.mark.parametrize('shape_in,size', [((2, 32, 32, 5), (128, 128)), ((1, 128, 128, 3), (512, 512)), ((2, 1, 1, 7), (32, 32)),
                                            ((1, 33, 33, 4), (129, 129)), ((1, 64, 128, 3), (256, 512)), ((1, 7, 5, 8), (25, 19)),
                                            ((1, 16, 16, 2), (7, 9))])
def test_resize_bilinear_equals_torch_half_pixel(shape_in, size):
    x = np.random.default_rng(0).standard_normal(shape_in).astype(np.float32)
    ours = R.resize_bilinear(x, size)
    ref = F.interpolate(_nchw(x), size=size, mode='bilinear', align_corners=False).permute(0, 2, 3, 1).numpy()
    assert np.abs(ours - ref).max() < 2e-6


def test_resize_1x1_is_broadcast():
    x = np.random.default_rng(1).standard_normal((2, 1, 1, 9)).astype(np.float32)
    assert np.array_equal(R.resize_bilinear(x, (8, 6)), np.broadcast_to(x, (2, 8, 6, 9)))


@pytest.mark.parametrize('rate', [1, 3, 6, 12, 18, 36])
def test_depthwise_equals_torch(rate):
    rng = np.random.default_rng(rate)
    x = rng.standard_normal((2, 20, 17, 6)).astype(np.float32)
    k = rng.standard_normal((3, 3, 6, 1)).astype(np.float32)
    ours = R.depthwise3x3(x, k, rate)
    kt = torch.from_numpy(k).permute(2, 3, 0, 1).contiguous()
    ref = F.conv2d(_nchw(x), kt, None, 1, rate, rate, groups=6).permute(0, 2, 3, 1).numpy()
    assert np.abs(ours - ref).max() < 1e-5


def test_conv1x1_bn_relu_equal_torch():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 5, 4, 24)).astype(np.float32)
    k = rng.standard_normal((1, 1, 24, 10)).astype(np.float32)
    b = rng.standard_normal(10).astype(np.float32)
    ref = F.conv2d(_nchw(x), torch.from_numpy(k).permute(3, 2, 0, 1).contiguous(), torch.from_numpy(b)).permute(0, 2, 3, 1).numpy()
    assert np.abs(R.conv1x1(x, k, b) - ref).max() < 1e-5
    g, be, mu, var = (rng.uniform(0.5, 1.5, 10).astype(np.float32) for _ in range(4))
    y = R.bn_inference(ref, g, be, mu, var, 1e-5)
    yt = F.batch_norm(_nchw(ref), torch.from_numpy(mu), torch.from_numpy(var), torch.from_numpy(g), torch.from_numpy(be),
                      False, 0.0, 1e-5).permute(0, 2, 3, 1).numpy()
    assert np.abs(y - yt).max() < 1e-5
    s, t = R.bn_fold(g, be, mu, var, 1e-5)
    assert np.abs(ref * s + t - y).max() < 1e-6


def test_bf16_round_equals_torch():
    x = (np.random.default_rng(5).standard_normal(100000) * 1000).astype(np.float32)
    assert np.array_equal(R.bf16_round(x), torch.from_numpy(x).bfloat16().float().numpy())
    bits = R.to_bf16_bits(x)
    assert np.array_equal(R.from_bf16_bits(bits), R.bf16_round(x))


def test_argmax_first_max_wins():
    x = np.zeros((1, 2, 2, 5), np.float32)
    x[0, 0, 0, [1, 3]] = 2.0
    x[0, 1, 1, [4, 2]] = 7.0
    lab = R.argmax_labels(x)
    assert lab[0, 0, 0] == 1 and lab[0, 1, 1] == 2 and lab[0, 0, 1] == 0


def test_miou_and_confusion_match_reference_formulae():
    gt = np.array([[0, 0, 1, 1], [2, 2, 255, 255]])
    pr = np.array([[0, 1, 1, 1], [2, 0, 0, 0]])
    # per-image mIOU over the labels present in gt (metrics.py:10-17): classes 0,1,2,255
    assert R.mIOU(gt, pr) == round((1 / 5 + 2 / 3 + 1 / 2 + 0) / 4, 2)
    cm = R.generate_matrix(gt, pr, 3)
    assert cm.sum() == 6 and cm[0, 0] == 1 and cm[0, 1] == 1 and cm[2, 0] == 1


@pytest.mark.parametrize('name', HEAD_CASES)
def test_oracle_reproduces_golden(name):
    """The committed vectors pin the oracle (regression) in both numeric modes."""
    cfg, W, feat, skip, z = load_case(name)
    for mode in ('fp32', 'bf16'):
        t = R.head_forward(feat, skip, W, cfg, mode)
        assert rel_err(t['logits'], z['logits_' + mode]) < 1e-5
        assert (t['labels'] == z['labels_' + mode]).mean() > 0.9995


@pytest.mark.parametrize('name', ['head_small_full', 'head_small_lite', 'head_odd_size'])
def test_numpy_oracle_equals_torch_backend(name):
    cfg, W, feat, skip, _ = load_case(name)
    a = R.head_forward(feat, skip, W, cfg, 'fp32')
    b = R.head_forward_torch(feat, skip, W, cfg, 'fp32')
    assert rel_err(a['logits'], b['logits'].numpy()) < 1e-5
    assert rel_err(a['logits_full'], b['logits_full'].numpy()) < 1e-5
    assert (a['labels'] == b['labels'].numpy()).mean() > 0.999


def test_bf16_mode_is_close_to_fp32_mode():
    cfg, W, feat, skip, z = load_case('head_small_full')
    assert rel_err(z['logits_bf16'], z['logits_fp32']) < 1e-2          # north_star: 1e-2 relative in bf16


def test_image_pool_branch_is_a_per_image_bias():
    """The fusion the CUDA path relies on: concat([b4, rest]) @ Wproj == rest @ Wproj[256:] + b4 @ Wproj[:256]."""
    rng = np.random.default_rng(9)
    b4 = rng.standard_normal((2, 1, 1, 256)).astype(np.float32)
    rest = rng.standard_normal((2, 4, 4, 1024)).astype(np.float32)
    k = rng.standard_normal((1, 1, 1280, 256)).astype(np.float32) * 0.05
    full = R.conv1x1(np.concatenate([R.resize_bilinear(b4, (4, 4)), rest], -1), k)
    split = R.conv1x1(rest, k[:, :, 256:]) + R.conv1x1(b4, k[:, :, :256])
    assert np.abs(full - split).max() < 2e-4


# ---- pins produced by EXECUTING the reference's own TensorFlow-free functions (tests/golden/make_ref_pins.py) ----------
def _pins():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_pins.npz'))


@pytest.mark.parametrize('k', [0, 1, 2, 3])
def test_metrics_reproduce_the_reference_functions(k):
    """deeplabv3p/metrics.py::mIOU and eval.py::generate_matrix, incl. missing classes and the ignore label 255."""
    z = _pins()
    gt, pr, nc = z['gt_%d' % k], z['pred_%d' % k], int(z['nc_%d' % k])
    assert R.mIOU(gt, pr) == float(z['miou_%d' % k])
    assert np.array_equal(R.generate_matrix(gt, pr, nc), z['confusion_%d' % k])


def test_preprocessing_reproduces_the_reference_functions():
    """common/data_utils.py: normalize_image, denormalize_image, preprocess_image (PIL bicubic), mask_resize (cv2 nearest)."""
    import os
    from PIL import Image
    z = _pins()
    assert np.array_equal(R.normalize_image(z['image_u8'].astype(np.float32)), z['normalize'])
    assert np.array_equal(R.denormalize_image(z['normalize']), z['denormalize'])
    assert np.array_equal(R.mask_resize(z['mask_in'], (27, 36)), z['mask_resize_27x36'])
    jpg = '/root/reference/example/2007_000039.jpg'
    if os.path.exists(jpg):                      # the example image itself lives in the reference tree (build container only)
        assert np.array_equal(R.preprocess_image(Image.open(jpg), (48, 64)), z['preprocess_2007_000039_48x64'])


PIL_CASES = [((64, 48), (32, 32)), ((37, 51), (64, 80)), ((375, 500), (512, 512)), ((512, 512), (512, 300)), ((96, 96), (96, 96)), ((9, 7), (40, 3)),
             ((300, 200), (31, 29)), ((33, 65), (16, 65))]


@pytest.mark.parametrize('hw_in,hw_out', PIL_CASES)
def test_pil_bicubic_restatement_equals_pillow(hw_in, hw_out):
    """preprocess_image's resize (common/data_utils.py:449) IS PIL's Image.resize(size, Image.BICUBIC): the integer restatement against
    Pillow itself, bit for bit, up- and down-scaling (antialiased), one axis unchanged, identity."""
    from PIL import Image
    rng = np.random.default_rng(hw_in[0] * 131 + hw_out[1])
    img = rng.integers(0, 256, hw_in + (3,)).astype(np.uint8)
    img[: hw_in[0] // 2] = (img[: hw_in[0] // 2] // 128) * 255          # saturated blocks: the overshoot of the cubic hits clip8
    ref = np.asarray(Image.fromarray(img).resize((hw_out[1], hw_out[0]), Image.BICUBIC))
    assert np.array_equal(R.pil_bicubic_resize(img, hw_out), ref)


def test_pil_bicubic_on_the_reference_example_image():
    import os
    from PIL import Image
    path = '/root/reference/example/2007_000039.jpg'
    if not os.path.exists(path):
        pytest.skip('reference example images are not on this machine')
    im = Image.open(path).convert('RGB')
    ref = np.asarray(im.resize((512, 512), Image.BICUBIC))
    assert np.array_equal(R.pil_bicubic_resize(np.asarray(im), (512, 512)), ref)
    assert np.array_equal(R.preprocess_image(im, (512, 512))[0], R.normalize_image(ref.astype('float32')))


def test_present_classes_follow_the_raster_order_of_the_native_postprocess():
    """inference/MNN/deeplabSegment.cpp:160-173 restated literally (per pixel: append a non-zero class on first sight)."""
    rng = np.random.default_rng(5)
    mask = rng.integers(0, 21, (40, 50)).astype(np.uint8)
    mask[:3] = 0
    mask[3, :10] = 7
    literal = []
    for v in mask.reshape(-1):
        if v != 0 and int(v) not in literal:
            literal.append(int(v))
    assert R.present_classes(mask) == literal and literal[0] == 7
    assert R.present_classes(np.zeros((4, 4), np.uint8)) == []


@pytest.mark.parametrize('n_in,n_out', [(64, 32), (37, 64), (375, 512), (500, 512), (512, 300), (9, 40), (7, 3), (300, 31), (2048, 512), (1, 5), (5, 1)])
def test_library_bicubic_coefficients_equal_the_restatement(n_in, n_out):
    """The host half of dlv3p_op_resize_bicubic_u8 (double arithmetic, 22-bit fixed point) needs no GPU: its tables against the oracle's
    restatement of Pillow's precompute_coeffs / normalize_coeffs_8bpc, integer for integer."""
    from dlv3p_b200 import ffi
    bounds, kk = ffi.pil_bicubic_coeffs(n_in, n_out)
    rb, rk = R._pil_bicubic_coeffs(n_in, n_out)
    assert np.array_equal(bounds, np.asarray(rb, np.int32))
    assert kk.shape == rk.shape and np.array_equal(kk, rk.astype(np.int32))
