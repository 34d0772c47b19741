"""CPU: the numpy twins of the kernels' dropout generators (tests/masks_np.py) -- known-answer test of Philox4x32-10, mask
statistics, determinism.  (Bit-exactness against the device generators is checked by the GPU tests.)"""
import numpy as np

from tests import masks_np as MK


def test_philox4x32_10_known_answers():
    # Random123 kat_vectors: philox4x32-10, counter / key all zero and all ones
    r = MK.philox(0, 0, 0, 0, 0, 0, 10)
    assert [int(v) for v in r] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    r = MK.philox(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 10)
    assert [int(v) for v in r] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]


def test_mask_rates_scales_and_determinism():
    cfg = dict(regularization=0.1, seq_len=10, hidden_dim=50, tokens_mlp_dim=20, channels_mlp_dim=50, num_blocks=2)
    for fn in (MK.mlp_generic_masks, MK.mlp_warp_masks, MK.mlp_tc5_masks):
        a, b, c = fn(cfg, 64, 123, 0), fn(cfg, 64, 123, 0), fn(cfg, 64, 123, 1)
        assert sorted(a) == sorted(b) and len(a) == 8
        for k in a:
            assert np.array_equal(a[k], b[k]) and not np.array_equal(a[k], c[k])
            vals = np.unique(a[k])
            assert len(vals) == 2 and vals[0] == 0 and abs(vals[1] - 1 / 0.9) < 1e-6
            assert abs(float((a[k] == 0).mean()) - 0.1) < 0.02
        assert a["Mixer_Block.0.mlp_block_token_mixing.reg1"].shape == (64, 50, 20)
        assert a["Mixer_Block.1.mlp_block_channel_mixing.reg2"].shape == (64, 10, 50)
    m = MK.conv_masks(dict(regularization=0.25, conv_nChan=2, in_nTP=10, dimPosEmb=50, num_blocks=1), 8, 5)
    assert m["Mixer_Block.0.conv1.reg"].shape == (8, 2, 10, 50) and abs(float((m["Mixer_Block.0.conv2.reg"] == 0).mean()) - 0.25) < 0.03
