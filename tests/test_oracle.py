"""CPU suite: the oracle against what pins it: parameter count, smp key scheme, torchvision's
real ResNet for the encoder, closed-form SSIM/noising cases, and the committed golden vectors."""
import math
import os

import numpy as np
import pytest
import torch
import torchvision

import oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "d3f_golden.npz")


def test_parameter_and_buffer_counts():
    m = oracle.Unet()
    assert sum(p.numel() for p in m.parameters()) == 24436659
    assert sum(p.numel() for p in oracle.Unet(classes=1).parameters()) == 24436369      # smp's published Unet-resnet34
    assert sum(b.numel() for b in m.buffers()) == 19054
    assert sum(1 for x in m.modules() if isinstance(x, torch.nn.Conv2d)) == 47
    assert sum(1 for x in m.modules() if isinstance(x, torch.nn.BatchNorm2d)) == 46


def test_state_dict_key_scheme():
    keys = set(oracle.Unet().state_dict())
    for k in ("encoder.conv1.weight", "encoder.bn1.running_var", "encoder.layer1.0.conv1.weight",
              "encoder.layer2.0.downsample.0.weight", "encoder.layer2.0.downsample.1.num_batches_tracked",
              "encoder.layer4.2.bn2.bias", "decoder.blocks.0.conv1.0.weight", "decoder.blocks.4.conv2.1.running_mean",
              "segmentation_head.0.weight", "segmentation_head.0.bias"):
        assert k in keys, k
    assert not any(k.startswith("encoder.fc") for k in keys)
    assert oracle.Unet().state_dict()["decoder.blocks.0.conv1.0.weight"].shape == (256, 768, 3, 3)
    assert oracle.Unet().state_dict()["decoder.blocks.4.conv1.0.weight"].shape == (16, 32, 3, 3)


def test_encoder_is_torchvision_resnet34():
    """Pin: the oracle's encoder features equal torchvision.models.resnet34's own layers on the same weights."""
    torch.manual_seed(0)
    m = oracle.Unet().eval()
    tv = torchvision.models.resnet34(weights=None).eval()
    sd = {k[len("encoder."):]: v for k, v in m.state_dict().items() if k.startswith("encoder.")}
    missing, unexpected = tv.load_state_dict(sd, strict=False)
    assert set(missing) == {"fc.weight", "fc.bias"} and not unexpected
    x = torch.randn(2, 3, 64, 64)
    with torch.no_grad():
        f = m.encoder(x)
        t = tv.relu(tv.bn1(tv.conv1(x)))
        assert torch.equal(f[1], t)
        t = tv.layer1(tv.maxpool(t))
        assert torch.equal(f[2], t)
        t = tv.layer4(tv.layer3(tv.layer2(t)))
        assert torch.equal(f[5], t)
    assert [tuple(a.shape[1:]) for a in f] == [(3, 64, 64), (64, 32, 32), (64, 16, 16), (128, 8, 8), (256, 4, 4), (512, 2, 2)]


def test_forward_shape_and_divisibility():
    m = oracle.Unet().eval()
    with torch.no_grad():
        assert m(torch.randn(1, 3, 32, 64)).shape == (1, 3, 32, 64)
    with pytest.raises(RuntimeError):
        m(torch.randn(1, 3, 48, 64))


def test_noise_ratio_distribution_and_formula():
    lam = 5.0
    y = torch.tensor([0.0, 0.5, 1.0 - 1e-7]).view(3, 1, 1, 1)
    r = oracle.sample_noise_ratio(y, lam)
    c = math.exp(-lam)
    assert abs(r[0].item() - 1.0) < 1e-6                                    # y = 0 -> r = 1
    assert abs(r[1].item() - (1 / lam) * math.log(1 / (0.5 * (1 - c) + c))) < 1e-7
    assert r[2].item() < 1e-6                                               # y -> 1 -> r -> 0
    g = torch.Generator().manual_seed(0)
    r = oracle.sample_noise_ratio(torch.rand(200000, 1, 1, 1, generator=g), lam)
    assert (r > 0).all() and (r <= 1).all()
    mean_theory = 1 / lam - math.exp(-lam) / (1 - math.exp(-lam))           # truncated exponential on (0, 1]
    assert abs(r.mean().item() - mean_theory) < 2e-3


def test_blend_is_variance_preserving():
    x = torch.randn(4, 3, 8, 8)
    n = torch.randn(4, 3, 8, 8)
    r = torch.tensor([0.0, 0.25, 0.5, 1.0]).view(4, 1, 1, 1)
    out = oracle.blend_noise(x, n, r)
    assert torch.equal(out[0], x[0]) and torch.equal(out[3], n[3])
    assert torch.allclose(out[1], math.sqrt(0.75) * x[1] + 0.5 * n[1], atol=1e-6)


def test_ssim_identities():
    g = torch.Generator().manual_seed(0)
    x = torch.rand(2, 3, 32, 32, generator=g)
    assert abs(oracle.ssim(x, x).item() - 1.0) < 1e-6
    assert oracle.ssim(x, 1 - x).item() < 0.2
    c1 = torch.full((1, 3, 16, 16), 0.2)
    c2 = torch.full((1, 3, 16, 16), 0.8)
    want = (2 * 0.2 * 0.8 + 1e-4) / (0.04 + 0.64 + 1e-4)                    # constant images: cs = 1, luminance only
    assert abs(oracle.ssim(c1, c2).item() - want) < 1e-5
    crit = oracle.MseStructuralSimilarityLoss(-1.0, 1.0)
    assert abs(crit(x * 2 - 1, x * 2 - 1).item()) < 1e-6


def test_sampler_degenerate_and_last_step():
    m = oracle.Unet().eval()
    x = torch.randn(1, 3, 32, 32)
    with torch.no_grad():
        assert torch.equal(oracle.sample_loop(m, x, 1, r_start=1.0), m(x))  # one step from r=1 returns x0_hat
    k = oracle.sampler.posterior_coeffs(0.6, 0.3, 0.0)                       # DDIM: deterministic
    assert k[2] == 0.0 and abs(k[0] - math.sqrt(0.3 / 0.6)) < 1e-12
    k = oracle.sampler.posterior_coeffs(0.6, 0.3, 1.0)
    var = k[0] ** 2 * 0.6 + k[2] ** 2                                        # noise variance carried to x_prev
    assert abs(var - 0.3) < 1e-12


def test_ema_restatement():
    lin = torch.nn.Linear(4, 4)
    ema = oracle.EMA(lin, beta=0.9999, update_every=1, include_online_model=False)
    assert "online_model" not in dict(ema.named_children())
    for step in range(105):
        with torch.no_grad():
            lin.weight.add_(1.0)
        ema.update()
        if step <= 100:
            assert torch.equal(ema.ema_model.weight, lin.weight)             # copies during warm-up
    assert not torch.equal(ema.ema_model.weight, lin.weight)
    assert 0.0 < ema.get_current_decay() < 0.9999


def test_golden_vectors():
    gold = {k: torch.from_numpy(v) for k, v in np.load(GOLD).items()}
    torch.manual_seed(0)
    model = oracle.Unet()
    r = oracle.sample_noise_ratio(gold["y"], 5.0)
    assert torch.allclose(r, gold["r"], atol=1e-7)
    noisy = oracle.blend_noise(gold["x"], gold["noise"], r)
    assert torch.allclose(noisy, gold["noisy"], atol=1e-6)
    model.train()
    pred = model(gold["noisy"])
    assert torch.allclose(pred, gold["pred_train"], atol=2e-5, rtol=1e-4)
    loss = oracle.MseStructuralSimilarityLoss(-1.0, 1.0)(pred, gold["x"])
    assert abs(loss.item() - gold["loss"].item()) < 1e-5
    loss.backward()
    g = model.segmentation_head[0].weight.grad
    assert (g - gold["head_w_grad"]).norm() / gold["head_w_grad"].norm() < 1e-4
    assert torch.allclose(model.encoder.bn1.running_mean, gold["bn1_running_mean"], atol=1e-6)
    model.eval()
    with torch.no_grad():
        assert torch.allclose(model(gold["noisy"]), gold["pred_eval"], atol=2e-5, rtol=1e-4)
    assert abs(oracle.ssim(gold["x"].clamp(0, 1), gold["noisy"].clamp(0, 1)).item() - gold["ssim_xn"].item()) < 1e-6


def test_frame_conversion_known_answers():
    """oracle.frames restates d3f/train_deep_fake/lit_module.py:272-300: BGR uint8 HWC <-> normalised RGB fp32 CHW."""
    import numpy as np
    frame = np.zeros((1, 2, 2, 3), dtype=np.uint8)
    frame[0, 0, 0] = (10, 20, 30)            # B, G, R
    frame[0, 1, 1] = (255, 0, 128)
    t = oracle.cv2_to_tensor_normalised(frame, [0.5, 0.5, 0.5], [0.5, 0.5, 0.5])
    assert t.shape == (1, 3, 2, 2) and t.dtype == torch.float32
    f32 = np.float32
    assert t[0, 0, 0, 0].item() == (f32(30) - f32(127.5)) / f32(127.5)       # R plane first
    assert t[0, 2, 0, 0].item() == (f32(10) - f32(127.5)) / f32(127.5)
    assert t[0, 0, 1, 1].item() == (f32(128) - f32(127.5)) / f32(127.5)
    assert t[0, 1, 0, 1].item() == -1.0 and t[0, 2, 1, 1].item() == 1.0
    # back: truncation toward zero, clamp, channel reversal
    x = torch.tensor([[-1.5, -1.0, 0.0, 0.999], [1.0, 1.7, 0.2, -0.004]]).reshape(1, 1, 2, 4).repeat(1, 3, 1, 1)
    x[0, 2] += 0.01
    out = oracle.tensor_cv2_to_denormalised(x, [0.5, 0.5, 0.5], [0.5, 0.5, 0.5])
    assert out.shape == (1, 2, 4, 3) and out.dtype == np.uint8
    assert out[0, 0, :, 2].tolist() == [0, 0, 127, 254] and out[0, 1, :, 2].tolist() == [255, 255, 153, 126]   # R from plane 0
    assert out[0, 0, 2, 0] == 128                                             # B from plane 2 (0.01 * 127.5 + 127.5 = 128.775)
    # the reference's own round trip is NOT the identity for every level (fp32 rounding + truncation): pin the count
    levels = np.arange(256, dtype=np.uint8).reshape(1, 16, 16, 1).repeat(3, axis=3)
    back = oracle.tensor_cv2_to_denormalised(oracle.cv2_to_tensor_normalised(levels, [0.5] * 3, [0.5] * 3), [0.5] * 3, [0.5] * 3)
    assert int((back != levels).sum()) == int((back.astype(int) - levels.astype(int) == -1).sum())   # only ever one level low


def test_affine_augmentation_restatement():
    """oracle.augment (kornia RandomAffine restated, d3f/train_denoiser/lit_module.py:55-65): parameter ranges, the
    centre-anchored map, and the bilinear / zero-padding warp pinned to torch's own grid_sample on the same inverse maps."""
    B, H, W = 6, 32, 48
    g = torch.Generator().manual_seed(3)
    angle, tx, ty, sc = oracle.sample_affine_params(B, H, W, generator=g)
    assert (angle.abs() <= 15).all() and (tx.abs() <= 0.2 * W).all() and (ty.abs() <= 0.2 * H).all()
    assert (sc >= 0.8).all() and (sc <= 1.2).all()
    M, Minv = oracle.affine_matrices(angle, tx, ty, sc, H, W)
    centre = torch.tensor([(W - 1) / 2, (H - 1) / 2, 1.0], dtype=torch.float64)
    moved = M @ centre                                       # the centre moves by exactly the translation
    assert torch.allclose(moved[:, 0] - centre[0], tx) and torch.allclose(moved[:, 1] - centre[1], ty)
    assert torch.allclose(torch.linalg.det(M[:, :2, :2]), sc ** 2)
    x = torch.randn(B, 3, H, W)
    out = oracle.warp_affine_bilinear(x, Minv[:, :2])
    oy, ox = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    sx = Minv[:, 0, 0, None, None] * ox + Minv[:, 0, 1, None, None] * oy + Minv[:, 0, 2, None, None]
    sy = Minv[:, 1, 0, None, None] * ox + Minv[:, 1, 1, None, None] * oy + Minv[:, 1, 2, None, None]
    grid = torch.stack([2 * sx / (W - 1) - 1, 2 * sy / (H - 1) - 1], dim=-1).float()
    ref = torch.nn.functional.grid_sample(x, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    assert (out - ref).abs().max() < 1e-4
    ident = torch.eye(3, dtype=torch.float64).repeat(B, 1, 1)
    assert torch.equal(oracle.warp_affine_bilinear(x, ident[:, :2]), x)
    # the product's closed-form inverse maps (host arithmetic, no device needed) are the oracle's matrix inverses
    from denoising_diffusion_deep_fake_b200.functional import random_affine_inverse_maps
    m = random_affine_inverse_maps(B, H, W, generator=torch.Generator().manual_seed(3))
    assert (m.view(B, 2, 3).double() - Minv[:, :2]).abs().max() < 1e-5


GOLD_F34 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "d3f_golden_f3_f4.npz")


def test_golden_vectors_rows_f3_f4():
    """The oracle reproduces the committed frame-conversion / affine-warp fixture (tests/golden/make_golden.py --rows-f3-f4)."""
    gold = dict(np.load(GOLD_F34))
    mean, std = gold["mean"].tolist(), gold["std"].tolist()
    assert torch.equal(oracle.cv2_to_tensor_normalised(gold["frames"], mean, std), torch.from_numpy(gold["tensor"]))
    assert np.array_equal(oracle.tensor_cv2_to_denormalised(torch.from_numpy(gold["net_out"]), mean, std), gold["frames_back"])
    _, minv = oracle.affine_matrices(*(torch.from_numpy(gold[k]) for k in ("aff_angle", "aff_tx", "aff_ty", "aff_scale")), 16, 24)
    assert np.allclose(minv[:, :2].numpy(), gold["aff_minv"], atol=1e-12)
    warped = oracle.warp_affine_bilinear(torch.from_numpy(gold["aff_x"]), torch.from_numpy(gold["aff_minv"]))
    assert torch.allclose(warped, torch.from_numpy(gold["aff_warped"]), atol=1e-6)
