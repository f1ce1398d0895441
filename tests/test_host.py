"""CPU tests of the host-side mirror (head.py) against the oracle's statement of the reference contract."""
import numpy as np
import pytest

import dlv3p_b200
from dlv3p_b200 import ffi, head
from oracle import head_ref as R


@pytest.mark.parametrize('lite,decoder,Cin,Cs,NC', [(False, True, 2048, 256, 21), (False, True, 320, 24, 19),
                                                     (True, False, 160, 24, 21), (True, True, 160, 24, 21)])
def test_weight_inventory_matches_reference_order(lite, decoder, Cin, Cs, NC):
    """Layer names / creation order / shapes of SURVEY §8(b) (load_weights(by_name=False), model.py:103)."""
    cfg = R.HeadConfig(B=1, H=128, W=128, OS=16, Cin=Cin, Cskip=Cs, NC=NC, lite=lite, decoder=decoder)
    stages = ffi.STAGE_ASPP | ffi.STAGE_TAIL | (ffi.STAGE_DECODER if decoder else 0)
    c = ffi.Context(device=-1, B=1, H=128, W=128, OS=16, Cin=Cin, Cskip=Cs if decoder else 0, NC=NC,
                    variant=1 if lite else 0, stages=stages)
    assert c.weight_specs() == R.weight_specs(cfg)


def test_head_param_counts_match_survey():
    def count(cfg):
        return sum(int(np.prod(s)) for _, _, s in R.weight_specs(cfg))
    # Xception 3 205 701 ; MobileNetV2 915 333 ; MNv3-L Lite 221 461
    assert count(R.HeadConfig(1, 512, 512, 16, 2048, 256, 21)) == 3205701
    assert count(R.HeadConfig(1, 512, 512, 16, 320, 24, 21)) == 915333
    assert count(R.HeadConfig(1, 512, 512, 16, 160, 24, 21, lite=True, decoder=False)) == 221461


def test_atrous_rates_and_errors():
    assert head.atrous_rates(16) == (6, 12, 18) and head.atrous_rates(8) == (12, 24, 36) and head.atrous_rates(32) == (3, 6, 9)
    with pytest.raises(ValueError):
        head.atrous_rates(4)
    assert R.atrous_rates(8) == head.atrous_rates(8)


def test_model_type_table():
    # same 18 keys as deeplab_model_map (model.py:23-48)
    assert len(head.MODEL_TYPES) == 18
    assert 'xception' in head.MODEL_TYPES and 'xception_lite' not in head.MODEL_TYPES
    with pytest.raises(ValueError):
        head.get_deeplabv3p_head('mobilenetv2lite', 21, (512, 512), 16)      # deeplab.py:32's typo'd default is not a key


def test_bf16_helpers_match_oracle_rounding():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(10000).astype(np.float32) * 50, np.float32([0, -0.0, 1e-30, 3.0e38, 1.00390625, 1.01171875])])
    assert np.array_equal(ffi.f32_to_bf16_bits(x), R.to_bf16_bits(x))
    assert np.array_equal(ffi.bf16_bits_to_f32(ffi.f32_to_bf16_bits(x)), R.bf16_round(x))
