"""q_sample / posterior_step parity (-m gpu) against the oracle formulas
(d3f/train_denoiser/lit_module.py:128-153).  fp32, tolerance 1e-6."""
import pytest
import torch

import oracle
import denoising_diffusion_deep_fake_b200 as d3
from gpu_harness import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("lam", [3.0, 5.0, 8.0])
@pytest.mark.parametrize("shape", [(8, 3, 64, 64), (1, 3, 32, 32), (5, 3, 128, 96)])
def test_q_sample_given_noise(lam, shape):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(shape, generator=g).clamp(-1, 1)
    noise = torch.randn(shape, generator=g)
    y = torch.rand(shape[0], 1, 1, 1, generator=g)
    r_ref = oracle.sample_noise_ratio(y, lam)
    ref = oracle.blend_noise(x, noise, r_ref)
    out, n_used, r = d3.q_sample(x.to(DEV), lam, noise=noise.to(DEV), y=y.to(DEV), return_aux=True)
    assert rel_err(out.cpu(), ref) < 1e-6
    assert rel_err(r.cpu(), r_ref) < 1e-6
    assert (r_ref > 0).all() and (r_ref <= 1).all()


def test_q_sample_philox_statistics():
    x = torch.zeros(64, 3, 64, 64, device=DEV)
    out, noise, r = d3.q_sample(x, 5.0, seed=123, offset=7, return_aux=True)
    # with x = 0: out = sqrt(r) * eps
    assert abs(noise.mean().item()) < 5e-3 and abs(noise.std().item() - 1.0) < 5e-3
    assert rel_err(out.cpu(), (torch.sqrt(r) * noise).cpu()) < 1e-6
    assert (r > 0).all() and (r <= 1).all()
    out2 = d3.q_sample(x, 5.0, seed=123, offset=7)
    assert torch.equal(out, out2)                     # deterministic given (seed, offset)
    out3 = d3.q_sample(x, 5.0, seed=124, offset=7)
    assert not torch.equal(out, out3)
    # fixed ratio variant (balance_training_images/lit_module.py:109-120)
    o4, n4, r4 = d3.q_sample(x + 1.0, 5.0, seed=1, fixed_r=0.7, return_aux=True)
    assert rel_err(o4.cpu(), ((0.3 ** 0.5) * 1.0 + (0.7 ** 0.5) * n4).cpu()) < 1e-6


@pytest.mark.parametrize("eta", [0.0, 1.0])
def test_posterior_step(eta):
    g = torch.Generator().manual_seed(1)
    shape = (4, 3, 64, 64)
    grid = oracle.noise_ratio_grid(10).tolist()
    for i in (0, 4, 9):
        x = torch.randn(shape, generator=g)
        x0 = torch.randn(shape, generator=g)
        z = torch.randn(shape, generator=g)
        ref = oracle.posterior_step(x, x0, grid[i], grid[i + 1], z, eta)
        xd = x.to(DEV).clone()
        d3.posterior_step_(xd, x0.to(DEV), grid[i], grid[i + 1], z=z.to(DEV), eta=eta)
        assert rel_err(xd.cpu(), ref) < 1e-6, i


@pytest.mark.parametrize("shape", [(8, 3, 64, 64), (2, 3, 32, 96), (3, 3, 128, 128)])
def test_fused_mse_ssim_loss(shape):
    """Fused loss kernel (value + dL/dprediction) vs the oracle's MseStructuralSimilarityLoss under autograd
    (d3f/loss_functions/structural_similarity_loss.py:14-26 + piqa SSIM).  fp32: value 1e-5, gradient 1e-4."""
    g = torch.Generator().manual_seed(3)
    target = torch.nn.functional.avg_pool2d(torch.randn(shape, generator=g), 5, 1, 2).clamp(-1, 1) * 1.5
    pred = (target + 0.4 * torch.randn(shape, generator=g)).requires_grad_(True)     # some values outside [-1, 1]
    ref = oracle.MseStructuralSimilarityLoss(-1.0, 1.0)(pred, target)
    ref.backward()
    crit = d3.MseStructuralSimilarityLoss(-1.0, 1.0)
    pd = pred.detach().to(DEV).requires_grad_(True)
    loss = crit(pd, target.to(DEV))
    (loss * 3.0).backward()
    assert abs(loss.item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item()))
    assert rel_err(pd.grad.cpu() / 3.0, pred.grad) < 1e-4
    # torch-composed fallback path of the same module agrees too (no-grad / non-fused inputs)
    with torch.no_grad():
        l2 = d3.loss.ssim(pd.clamp(-1, 1) * 0.5 + 0.5, target.to(DEV).clamp(-1, 1) * 0.5 + 0.5)
        assert abs(l2.item() - oracle.ssim(pred.detach().clamp(-1, 1) * 0.5 + 0.5, target.clamp(-1, 1) * 0.5 + 0.5).item()) < 1e-5


@pytest.mark.parametrize("n,betas,with_ema", [(1000003, (0.9, 0.999), False), (4096, (0.5, 0.999), True), (37, (0.9, 0.999), True)])
def test_fused_adam_matches_torch_optim(n, betas, with_ema):
    """d3fk_adam vs torch.optim.Adam (d3f/train_denoiser/lit_module.py:95, train_deep_fake/lit_module.py:116-120) and the
    ema_pytorch lerp, three steps, odd sizes exercise the scalar tail."""
    from denoising_diffusion_deep_fake_b200.functional import adam_step_
    g = torch.Generator().manual_seed(11)
    npad = (n + 3) // 4 * 4 + 4
    p0 = torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=0.02, betas=betas)
    flat = torch.zeros(npad, device=DEV)
    p = flat[:n]
    p.copy_(p0)
    m, v = torch.zeros(npad, device=DEV)[:n], torch.zeros(npad, device=DEV)[:n]
    ema = torch.zeros(npad, device=DEV)[:n] if with_ema else None
    ema_ref = torch.zeros(n)
    for step in range(1, 4):
        grad = torch.randn(n, generator=g)
        ref.grad = grad.clone()
        opt.step()
        adam_step_(p, grad.to(DEV), m, v, 0.02, betas[0], betas[1], 1e-8, step, ema=ema, ema_decay=0.9)
        ema_ref.lerp_(ref.detach(), 1 - 0.9)
    assert rel_err(p.cpu(), ref.detach()) < 1e-6
    if with_ema:
        assert rel_err(ema.cpu(), ema_ref) < 1e-6


FRAME_NORMS = [([0.5, 0.5, 0.5], [0.5, 0.5, 0.5]), ([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])]


@pytest.mark.parametrize("mean,std", FRAME_NORMS)
@pytest.mark.parametrize("shape", [(3, 32, 64), (1, 128, 128), (2, 6, 10), (5, 448, 448)])
def test_frames_to_tensor_bit_exact(mean, std, shape):
    """d3fk_frames_to_tensor vs the reference arithmetic (d3f/train_deep_fake/lit_module.py:272-283): bit-exact."""
    import numpy as np
    from denoising_diffusion_deep_fake_b200.functional import frames_to_tensor
    N, H, W = shape
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 256, size=(N, H, W, 3), dtype=np.uint8)
    frames[0, 0, :4] = [[0, 0, 0], [255, 255, 255], [0, 128, 255], [1, 2, 3]]
    ref = oracle.cv2_to_tensor_normalised(frames, mean, std)
    out = frames_to_tensor(torch.from_numpy(frames).to(DEV), mean, std)
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert torch.equal(out.cpu(), ref)
    one = frames_to_tensor(torch.from_numpy(frames[0]).to(DEV), mean, std)        # a single [H,W,3] frame
    assert torch.equal(one.cpu(), ref[:1])


@pytest.mark.parametrize("mean,std", FRAME_NORMS)
@pytest.mark.parametrize("shape", [(3, 32, 64), (1, 128, 128), (2, 6, 10), (5, 448, 448)])
def test_tensor_to_frames_bit_exact(mean, std, shape):
    """d3fk_tensor_to_frames vs the reference arithmetic (:285-300): bit-exact bytes, including truncation toward zero,
    the clamp, values far out of range and exact level boundaries; the input tensor is not modified.  (Beyond +-2^31 the
    reference's `tensor.int()` is device-dependent — x86 yields INT_MIN, CUDA saturates; d3fk saturates like the CUDA
    device the reference runs this on — so the test stays inside the int32 range.)"""
    import numpy as np
    from denoising_diffusion_deep_fake_b200.functional import tensor_to_frames
    N, H, W = shape
    g = torch.Generator().manual_seed(1)
    t = torch.randn(N, 3, H, W, generator=g) * 1.2
    flat = t.view(-1)
    flat[:8] = torch.tensor([-1.0, 1.0, 0.0, 1e6, -1e6, 16000000.0, -0.0039215689, 1.0000001])
    levels = torch.arange(256, dtype=torch.float32)
    flat[8:8 + 256] = (levels - 127.5) / 127.5                   # exact level boundaries of the 0.5 / 0.5 normalisation
    ref = oracle.tensor_cv2_to_denormalised(t, mean, std)
    t_dev = t.to(DEV)
    keep = t_dev.clone()
    out = tensor_to_frames(t_dev, mean, std)
    assert out.dtype == torch.uint8 and tuple(out.shape) == (N, H, W, 3)
    assert np.array_equal(out.cpu().numpy(), ref)
    assert torch.equal(t_dev, keep)


def test_frame_round_trip_and_errors():
    """frames -> tensor -> frames equals the reference's own round trip (which loses one level on some values), at a full
    video-frame size; malformed inputs fail loudly."""
    import numpy as np
    from denoising_diffusion_deep_fake_b200.functional import frames_to_tensor, tensor_to_frames
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 256, size=(4, 448, 448, 3), dtype=np.uint8)
    mean, std = [0.5] * 3, [0.5] * 3
    ref = oracle.tensor_cv2_to_denormalised(oracle.cv2_to_tensor_normalised(frames, mean, std), mean, std)
    out = tensor_to_frames(frames_to_tensor(torch.from_numpy(frames).to(DEV), mean, std), mean, std)
    assert np.array_equal(out.cpu().numpy(), ref)
    assert int(np.abs(out.cpu().numpy().astype(int) - frames.astype(int)).max()) <= 1
    with pytest.raises(d3._lib.D3fkError):
        frames_to_tensor(torch.from_numpy(frames), mean, std)                     # CPU tensor: no CPU path
    with pytest.raises(RuntimeError):
        frames_to_tensor(torch.zeros(1, 8, 8, 3, device=DEV), mean, std)          # not uint8
    with pytest.raises(d3._lib.D3fkError):
        frames_to_tensor(torch.zeros(1, 3, 3, 3, dtype=torch.uint8, device=DEV), mean, std)   # H*W % 4 != 0
    with pytest.raises(d3._lib.D3fkError):
        tensor_to_frames(torch.zeros(1, 3, 8, 8, device=DEV), mean, [0.5, 0.0, 0.5])          # std == 0


def test_predict_fake_on_frames():
    """DeepFakeModule.predict_fake with the reference's signature (one uint8 BGR frame in, one out,
    d3f/train_deep_fake/lit_module.py:251-270) against the oracle chain normalise -> U-Net (eval) -> denormalise."""
    import numpy as np
    from denoising_diffusion_deep_fake_b200.train import DeepFakeModule
    torch.manual_seed(0)
    mod = DeepFakeModule(encoder_name="resnet34", learning_rate=1e-3, noise_exponential_sampling_lambda=3, mode="denoise",
                         mean_a=[0.5, 0.5, 0.5], std_a=[0.5, 0.5, 0.5], mean_b=[0.45, 0.5, 0.55], std_b=[0.4, 0.5, 0.6],
                         precision="fp32").to(DEV)
    ref = oracle.Unet()
    ref.load_state_dict(mod.model_a.state_dict())
    ref.eval()
    rng = np.random.default_rng(5)
    frame = rng.integers(0, 256, size=(64, 64, 3), dtype=np.uint8)
    fake = mod.predict_fake(frame, "a")
    assert isinstance(fake, np.ndarray) and fake.dtype == np.uint8 and fake.shape == frame.shape
    with torch.no_grad():
        x = oracle.cv2_to_tensor_normalised(frame[None], [0.45, 0.5, 0.55], [0.4, 0.5, 0.6])     # model a uses mean_b / std_b
        want = oracle.tensor_cv2_to_denormalised(ref(x), [0.45, 0.5, 0.55], [0.4, 0.5, 0.6])[0]
    diff = np.abs(fake.astype(int) - want.astype(int))
    assert diff.max() <= 1 and (diff != 0).mean() < 1e-2      # fp32 network within 1e-5: at most a level flips at a boundary
    batch = torch.from_numpy(np.stack([frame, frame[::-1].copy()])).to(DEV)
    out = mod.predict_fake(batch, "a")
    assert out.is_cuda and out.dtype == torch.uint8 and np.array_equal(out[0].cpu().numpy(), fake)


@pytest.mark.parametrize("shape", [(8, 3, 64, 64), (3, 3, 32, 96), (2, 3, 128, 128)])
def test_affine_q_sample(shape):
    """d3fk_affine_q_sample (d3f/train_denoiser/lit_module.py:113-115): the warp against the oracle's bilinear /
    zero-padding restatement on the same inverse maps, and the noising bit-identical to q_sample of the warped image."""
    from denoising_diffusion_deep_fake_b200.functional import affine_q_sample, random_affine_inverse_maps
    B, C, H, W = shape
    g = torch.Generator().manual_seed(2)
    x = torch.randn(shape, generator=g).clamp(-1, 1)
    m = random_affine_inverse_maps(B, H, W, generator=g)
    noise = torch.randn(shape, generator=g)
    y = torch.rand(B, 1, 1, 1, generator=g)
    aug_ref = oracle.warp_affine_bilinear(x, m.view(B, 2, 3))
    r_ref = oracle.sample_noise_ratio(y, 5.0)
    aug, noisy, r = affine_q_sample(x.to(DEV), m, 5.0, noise=noise.to(DEV), y=y.to(DEV), return_aux=True)
    # tolerance: the source coordinate is rounded differently (fma in the kernel): 1 ulp at ~100 px = 8e-6 px times an image
    # slope of up to 2 per pixel
    assert (aug.cpu() - aug_ref).abs().max() < 1e-4, (aug.cpu() - aug_ref).abs().max()
    assert rel_err(r.cpu(), r_ref) < 1e-6
    assert rel_err(noisy.cpu(), oracle.blend_noise(aug_ref, noise, r_ref)) < 2e-5
    # fused == two kernels, bit for bit, on both RNG paths
    assert torch.equal(noisy, d3.q_sample(aug, 5.0, noise=noise.to(DEV), y=y.to(DEV)))
    aug2, noisy2 = affine_q_sample(x.to(DEV), m, 5.0, seed=11, offset=4)
    assert torch.equal(aug2, aug) and torch.equal(noisy2, d3.q_sample(aug2, 5.0, seed=11, offset=4))
    # identity map: the image itself; a map that leaves the image: zeros
    ident = torch.tensor([1.0, 0, 0, 0, 1, 0]).repeat(B, 1)
    a_id, _ = affine_q_sample(x.to(DEV), ident, 5.0, seed=1)
    assert torch.equal(a_id.cpu(), x)
    far = torch.tensor([1.0, 0, 10.0 * W, 0, 1, 0]).repeat(B, 1)
    a_far, n_far, r_far = affine_q_sample(x.to(DEV), far, 5.0, noise=noise.to(DEV), y=y.to(DEV), return_aux=True)
    assert float(a_far.abs().max()) == 0.0 and rel_err(n_far.cpu(), torch.sqrt(r_ref) * noise) < 1e-6


def test_training_step_with_augmentation():
    """DenoiserModule(augment=True): the training step warps and noises in one kernel and regresses onto the WARPED image
    (the reference's order: augment, noise, predict, loss against the augmented image — lit_module.py:113-119)."""
    from denoising_diffusion_deep_fake_b200.train import DenoiserModule
    torch.manual_seed(0)
    mod = DenoiserModule(encoder_name="resnet34", learning_rate=1e-3, noise_exponential_sampling_lambda=5,
                         cosine_scheduler_max_epoch=10, precision="bf16", seed=5, augment=True).to(DEV).train()
    mod.configure_optimizers(fused=True)
    g = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn(8, 3, 64, 64, generator=g, device=DEV).clamp(-1, 1)
    losses = [float(mod.training_step(x)) for _ in range(12)]
    assert all(l == l and l < 10 for l in losses) and min(losses[-4:]) < losses[0]
    aug, noisy = mod.augment_and_blend(x)
    assert aug.shape == x.shape and not torch.equal(aug, x) and torch.isfinite(noisy).all()


def test_rows_f3_f4_against_committed_golden_vectors():
    """The CUDA kernels against tests/golden/d3f_golden_f3_f4.npz: frame bytes exactly, warp within coordinate rounding."""
    import os
    import numpy as np
    from denoising_diffusion_deep_fake_b200.functional import affine_q_sample, frames_to_tensor, tensor_to_frames
    gold = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "d3f_golden_f3_f4.npz")))
    mean, std = gold["mean"].tolist(), gold["std"].tolist()
    t = frames_to_tensor(torch.from_numpy(gold["frames"]).to(DEV), mean, std)
    assert torch.equal(t.cpu(), torch.from_numpy(gold["tensor"]))
    back = tensor_to_frames(torch.from_numpy(gold["net_out"]).to(DEV), mean, std)
    assert np.array_equal(back.cpu().numpy(), gold["frames_back"])
    x = torch.from_numpy(gold["aff_x"]).to(DEV)
    aug, _ = affine_q_sample(x, torch.from_numpy(gold["aff_minv"]).float(), 5.0, seed=1)
    assert (aug.cpu() - torch.from_numpy(gold["aff_warped"])).abs().max() < 1e-4
