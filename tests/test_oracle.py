"""The oracle (oracle/pangu_oracle.py) against the committed golden fixtures, which are outputs
of the unmodified reference run by oracle/make_golden.py -- plus, in the build container,
against the live reference modules.  CPU only."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import pangu_oracle as O
from tests.util import golden, sampled_rel_l2


def test_window_maps_bit_exact():
    g = golden("index_maps.npz")
    for tag, H in (("hi", 181), ("lo", 91)):
        for roll in (0, 1):
            ref = torch.from_numpy(g[f"{tag}.src.roll{roll}"].astype(np.int64))
            assert torch.equal(O.window_source_index(8, H, 24, bool(roll)), ref)


def test_shift_mask_bit_exact():
    g = golden("index_maps.npz")
    for tag, H in (("hi", 181), ("lo", 91)):
        m = O.shift_mask(8, H)
        bits = np.packbits((m != 0).numpy().reshape(-1))
        assert np.array_equal(bits, g[f"{tag}.mask_bits"])
        assert set(m.unique().tolist()) <= {0.0, -100.0}
    # pattern census: 90 unmasked types, 30 z-split + 3 h-split (half of the
    # pairs masked each), 1 type with both splits (three quarters masked)
    m = (O.shift_mask(8, 181) != 0).flatten(1).sum(1)
    vals, counts = torch.unique(m, return_counts=True)
    assert vals.tolist() == [0, 10368, 15552] and counts.tolist() == [90, 33, 1]


def test_position_index():
    g = golden("index_maps.npz")
    idx = O.position_index()
    assert np.array_equal(idx.numpy().astype(np.int16), g["position_index"])
    assert idx.min() == 0 and idx.max() == 3311 and idx.numel() == 20736


def test_partition_reverse_is_identity_on_real_tokens():
    for H in (181, 91):
        for roll in (False, True):
            src = O.window_source_index(8, H, 36, roll).reshape(-1)
            real = src[src >= 0]
            assert real.numel() == 8 * H * 36 and torch.equal(real.sort().values, torch.arange(8 * H * 36))
            assert (src < 0).sum() == 8 * 5 * 36


def test_blocks_against_golden():
    g = golden("blocks.npz")
    p = O.stress_weights(seed=7)
    for tag, dim, heads, H, pre in (("hi", 192, 6, 181, "layers.EarthSpecificLayer0.blocks.EarthSpecificBlock1."),
                                    ("lo", 384, 12, 91, "layers.EarthSpecificLayer1.blocks.EarthSpecificBlock1.")):
        x = torch.randn(1, 8 * H * 24, dim, generator=torch.Generator().manual_seed(11))
        for roll in (0, 1):
            y = O.earth_block(x, p, pre, 8, H, 24, heads, bool(roll))
            assert sampled_rel_l2(y, g, f"{tag}.roll{roll}") < 5e-6
            assert abs(float(y.double().norm()) / float(g[f"{tag}.roll{roll}.l2"]) - 1) < 1e-5
    x = torch.randn(1, 8 * 181 * 24, 192, generator=torch.Generator().manual_seed(12))
    assert sampled_rel_l2(O.down_sample(x, p, 8, 181, 24), g, "down") < 5e-6


def test_param_shapes_match_reference_layout():
    shapes = O.param_shapes()
    assert len(shapes) == 223
    assert sum(int(np.prod(s)) for _, s in shapes) == 276_659_936


def test_loss_and_norm_helpers():
    gen = torch.Generator().manual_seed(0)
    ou, os_ = torch.randn(1, 5, 13, 8, 8, generator=gen), torch.randn(1, 4, 8, 8, generator=gen)
    tu, ts = torch.randn(1, 5, 13, 8, 8, generator=gen), torch.randn(1, 4, 8, 8, generator=gen)
    crit = torch.nn.L1Loss(reduction="none")                               # models/pangu_sample.py:18
    wu = torch.FloatTensor(O.UPPER_WEIGHTS).view(1, 5, 1, 1, 1)
    ws = torch.FloatTensor(O.SURFACE_WEIGHTS).view(1, 4, 1, 1)
    want = torch.mean(crit(ou, tu) * wu) + torch.mean(crit(os_, ts) * ws) * 0.25
    assert torch.allclose(O.weighted_l1_loss(ou, os_, tu, ts), want, rtol=1e-6)
    stats = (torch.randn(4), 0.5 + torch.rand(4), torch.randn(13, 1, 1, 5), 0.5 + torch.rand(13, 1, 1, 5))
    ost = O.output_statistics(stats)
    nu, ns = O.norm_data(tu, ts, ost)
    bu, bs = O.norm_back_data(nu, ns, ost)
    assert torch.allclose(bu, tu, atol=1e-5) and torch.allclose(bs, ts, atol=1e-5)
    # level l of the data is normalised with row 12-l of the (13,1,1,5) input statistics
    assert torch.equal(ost[2][0, :, 3, 0, 0], stats[2][9, 0, 0, :])


def test_strip_forward_runs_and_is_deterministic():
    p = O.reference_like_weights(seed=0)
    up, sf, stats, maps, ch = O.synthetic_inputs(seed=1, lat=721, lon=96)
    a = O.forward(p, up, sf, stats, maps, ch)
    b = O.forward(p, up, sf, stats, maps, ch)
    assert a[0].shape == (1, 5, 13, 721, 96) and a[1].shape == (1, 4, 721, 96)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert torch.isfinite(a[0]).all() and 0.3 < float(a[0].std()) < 3.0


# ---------------------------------------------------------------------------------------
# live reference (build container only; the full-size cases take ~2 min each and are what
# oracle/make_golden.py asserts while writing the fixtures, so only a cheap case runs here)
# ---------------------------------------------------------------------------------------
@pytest.mark.reference
def test_block_against_live_reference():
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(os.path.dirname(here), "oracle", "timm_shim"))
    sys.path.insert(0, os.environ.get("PANGU_REFERENCE", "/root/reference"))
    from models import layers as RL
    p = O.stress_weights(seed=5)
    pre = "layers.EarthSpecificLayer1.blocks.EarthSpecificBlock3."
    blk = RL.EarthSpecificBlock(384, 0.0, 12, device="cpu").eval()
    blk.load_state_dict({k[len(pre):]: v for k, v in p.items() if k.startswith(pre)}, strict=True)
    x = torch.randn(1, 8 * 91 * 12, 384, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        for roll in (False, True):
            ref = blk(x, 8, 91, 12, roll)
            mine = O.earth_block(x, p, pre, 8, 91, 12, 12, roll)
            assert float((mine - ref).norm() / ref.norm()) < 2e-6
    for m in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        del sys.modules[m]


@pytest.mark.reference
def test_scores_against_live_reference():
    """oracle RMSE / ACC == era5_data/score.py (imported unmodified; it needs only numpy + torch)."""
    ref_root = os.environ.get("PANGU_REFERENCE", "/root/reference")
    sys.path.insert(0, ref_root)
    try:
        from era5_data import score
    finally:
        sys.path.remove(ref_root)
    g = torch.Generator().manual_seed(0)
    a, b = torch.randn(13, 721, 48, generator=g), torch.randn(13, 721, 48, generator=g)
    assert torch.allclose(O.weighted_rmse_channels(a, b), score.weighted_rmse_torch_channels(a, b), rtol=1e-5)
    assert torch.allclose(O.weighted_acc_channels(a, b), score.weighted_acc_torch_channels(a, b), rtol=1e-5, atol=1e-7)
    a4, b4 = a.unsqueeze(0), b.unsqueeze(0)
    assert torch.allclose(O.weighted_rmse_channels(a4, b4), score.weighted_rmse_torch_channels(a4, b4), rtol=1e-5)


@pytest.mark.reference
def test_block_gradients_against_live_reference():
    """Oracle autograd == reference autograd for one EarthSpecificBlock per resolution and roll state (asserted inside)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    try:
        import make_golden
    finally:
        sys.path.pop(0)
    make_golden.gen_block_grads()
