"""Pins the oracle to the REFERENCE'S OWN CODE, run here (build container only: /root/reference does not travel).

    python tests/golden/make_ref_golden.py        ->  tests/golden/ref_golden.npz

The reference's LitModules cannot be imported (pytorch_lightning / segmentation_models_pytorch / piqa / kornia /
ema_pytorch are not installed — SURVEY §8c), but the hot-path functions that are pure torch / numpy / cv2 can be
lifted out of their class bodies with `ast`, compiled unchanged, and executed on a stub `self`:

  d3f/train_denoiser/lit_module.py    :107-126  training_step (flow; model / criterion / augmentation are stand-ins)
                                      :128-139  blend_random_amount_of_noise_with_each_sample
                                      :141-153  sample_random_number_from_exponential_distribution
                                      :92-100   configure_optimizers (Adam + CosineAnnealingLR)
  d3f/train_deep_fake/lit_module.py   :168-181  training_denoise_step_for_one_model (flow)
                                      :183-206  training_swap_step_for_one_model (flow)
                                      :208-233  the two noising functions (duplicates)
                                      :113-125  configure_optimizers (two Adams, betas from hparams)
                                      :272-300  cv2_to_tensor_normalised / tensor_cv2_to_denormalised
  d3f/balance_training_images/lit_module.py :109-120  blend_fixed_amount_of_noise_with_each_sample
  d3f/loss_functions/structural_similarity_loss.py :14-26  forward / normalise_between_zero_and_one
                                      (piqa.SSIM is a stand-in: the oracle's restated SSIM — SSIM itself stays unpinned)

No reference SOURCE is copied into the repo: the functions are compiled from the files where they lie and only their
input / output tensors are stored.  tests/test_ref_pin.py asserts oracle == these outputs to 0 ulp."""
import ast
import math
import os
import sys
import types

import cv2
import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

REF = "/root/reference/d3f"


def extract(path, cls, names, extra_globals=None):
    """Compile the named methods of class `cls` in `path` (unchanged) and return them as plain functions."""
    tree = ast.parse(open(os.path.join(REF, path)).read())
    glob = {"torch": torch, "math": math, "np": np, "cv2": cv2, "nn": nn,
            "optimizers": torch.optim, "schedulers": torch.optim.lr_scheduler}
    glob.update(extra_globals or {})
    out = {}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name in names:
                    mod = ast.Module(body=[fn], type_ignores=[])
                    exec(compile(mod, os.path.join(REF, path), "exec"), glob)
                    out[fn.name] = glob[fn.name]
    missing = set(names) - set(out)
    assert not missing, (path, cls, missing)
    return out


class Stub:
    """Stand-in for a LightningModule instance: hparams namespace, device, no-op logging."""

    def __init__(self, fns, **hparams):
        self.hparams = types.SimpleNamespace(**hparams)
        self.device = torch.device("cpu")
        self.global_step = 0
        self.logged = {}
        self.image_logging_scheduler = types.SimpleNamespace(update_with_step_number=lambda step: None)
        for name, fn in fns.items():
            setattr(self, name, types.MethodType(fn, self))

    def log(self, name, value, *a, **k):
        self.logged[name] = value.detach().clone() if torch.is_tensor(value) else value

    def log_batch_as_image_grid(self, *a, **k):
        pass


def small_net(seed):
    """Deterministic stand-in for smp.Unet in the FLOW pins (the flow does not depend on what the model computes)."""
    torch.manual_seed(seed)
    return nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.BatchNorm2d(8), nn.ReLU(), nn.Conv2d(8, 3, 3, padding=1))


def main():
    out = {}
    noising = ("blend_random_amount_of_noise_with_each_sample", "sample_random_number_from_exponential_distribution")
    den = extract("train_denoiser/lit_module.py", "LitModule", noising + ("training_step", "configure_optimizers"))
    fake = extract("train_deep_fake/lit_module.py", "LitModule",
                   noising + ("training_denoise_step_for_one_model", "training_swap_step_for_one_model",
                              "configure_optimizers", "cv2_to_tensor_normalised", "tensor_cv2_to_denormalised",
                              "predict_fake_for_single_frame"))
    bal = extract("balance_training_images/lit_module.py", "LitModule", ("blend_fixed_amount_of_noise_with_each_sample",))
    crit_fns = extract("loss_functions/structural_similarity_loss.py", "MseStructuralSimilarityLoss",
                       ("forward", "normalise_between_zero_and_one"))

    # ---- a1 / a2: noising, every lambda the configs use (denoiser_config.yml:8, denoise_config.yml:11, swap_config.yml:8)
    g = torch.Generator().manual_seed(1234)
    x = torch.nn.functional.avg_pool2d(0.5 * torch.randn(5, 3, 16, 16, generator=g), 5, 1, 2).mul(2.5).clamp(-1, 1)
    out["q_x"] = x
    for lam in (3, 5, 8):
        s_den = Stub(den, noise_exponential_sampling_lambda=lam)
        s_fake = Stub(fake, noise_exponential_sampling_lambda=lam)
        torch.manual_seed(100 + lam)
        noisy = s_den.blend_random_amount_of_noise_with_each_sample(x)
        torch.manual_seed(100 + lam)
        noisy2 = s_fake.blend_random_amount_of_noise_with_each_sample(x)
        assert torch.equal(noisy, noisy2), "the two copies of the noising in the reference disagree"
        torch.manual_seed(200 + lam)
        r = s_den.sample_random_number_from_exponential_distribution(7, lam)
        out[f"q_noisy_lam{lam}"] = noisy
        out[f"q_r_lam{lam}"] = r
    # edge values of the inverse-CDF: y = 0 -> r = 1, y -> 1 -> r -> 0 (evaluated through the reference's own expression by
    # monkey-patching torch.rand for this one call)
    edge_y = torch.tensor([0.0, 1e-9, 0.25, 0.5, 0.999999, 1.0]).view(-1, 1, 1, 1)
    real_rand = torch.rand
    try:
        torch.rand = lambda size, device=None: edge_y.clone()
        out["q_edge_y"] = edge_y
        out["q_edge_r_lam5"] = Stub(den).sample_random_number_from_exponential_distribution(6, 5)
    finally:
        torch.rand = real_rand
    s_bal = Stub(bal, ratio_of_noise=0.7)
    torch.manual_seed(77)
    out["q_fixed07"] = s_bal.blend_fixed_amount_of_noise_with_each_sample(x)

    # ---- a5 wrapper: normalise + (mse + 1 - ssim) / 2 with the oracle's SSIM standing in for piqa.SSIM
    crit = Stub(crit_fns)
    crit.input_min_value, crit.input_max_value = -1.0, 1.0
    crit.ssim = oracle.ssim
    crit.mse = nn.MSELoss()
    pred = x * 1.4 + 0.2 * torch.randn(x.shape, generator=g)        # leaves [-1, 1]: exercises the clip
    out["crit_pred"], out["crit_target"] = pred, x
    out["crit_norm_pred"] = crit.normalise_between_zero_and_one(pred)
    out["crit_loss"] = crit.forward(pred, x)

    # ---- f4: frame conversion around predict_fake (real cv2.cvtColor), one frame at a time as the reference does
    rng = np.random.default_rng(7)
    frames = rng.integers(0, 256, size=(3, 10, 14, 3), dtype=np.uint8)
    frames[0, :2] = 0
    frames[0, 2:4] = 255
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    s = Stub(fake)
    mt, st = torch.tensor(mean), torch.tensor(std)
    tens = torch.cat([s.cv2_to_tensor_normalised(f, mt, st) for f in frames])
    net_out = tens * 1.3 + 0.3 * torch.randn(tens.shape, generator=g)   # pushes some pixels below 0 / above 255
    back = np.stack([s.tensor_cv2_to_denormalised(t.clone().unsqueeze(0), mt, st) for t in net_out])
    out.update(fr_frames=frames, fr_mean=np.array(mean, np.float32), fr_std=np.array(std, np.float32), fr_tensor=tens,
               fr_net_out=net_out, fr_back=back)
    # predict_fake_for_single_frame with a stand-in model: pre -> model -> post in one call
    net = small_net(5).eval()
    with torch.no_grad():
        out["fr_predict"] = np.stack([s.predict_fake_for_single_frame(f, net, mean, std) for f in frames])

    # ---- a6: optimiser construction and the per-epoch cosine schedule
    s = Stub(den, learning_rate=0.02, cosine_scheduler_max_epoch=10)
    s.model = nn.Linear(2, 2)
    (opt,), (sch,) = s.configure_optimizers()
    lrs = []
    for _ in range(12):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sch.step()
    d = opt.defaults
    out["opt_den_lrs"] = np.array(lrs)
    out["opt_den_cfg"] = np.array([d["lr"], d["betas"][0], d["betas"][1], d["eps"], d["weight_decay"], float(d["amsgrad"])])
    s = Stub(fake, learning_rate=0.02, cosine_scheduler_max_epoch=50, adam_b1=0.5, adam_b2=0.999)
    s.model_a, s.model_b = nn.Linear(2, 2), nn.Linear(2, 2)
    opts, schs = s.configure_optimizers()
    d = opts[1].defaults
    out["opt_fake_cfg"] = np.array([d["lr"], d["betas"][0], d["betas"][1], d["eps"], d["weight_decay"], float(d["amsgrad"]),
                                    len(opts), schs[0].T_max])

    # ---- step flows (a3-a5, a7) with stand-in networks: which tensor goes where, in which RNG order
    crit_mod = oracle.MseStructuralSimilarityLoss(-1.0, 1.0)
    s = Stub(den, noise_exponential_sampling_lambda=5)
    s.model = small_net(1).train()
    s.shared_augmentation_sequence = lambda im: im
    s.training_criterion = crit_mod
    torch.manual_seed(300)
    loss = s.training_step({"image": x}, 0)
    out["flow_denoiser_loss"] = loss.detach()

    s = Stub(fake, noise_exponential_sampling_lambda=3, mode="denoise")
    s.criterion = crit_mod
    real_model = small_net(2).train()
    torch.manual_seed(301)
    out["flow_fake_denoise_loss"] = s.training_denoise_step_for_one_model("a", x, real_model).detach()

    s = Stub(fake, noise_exponential_sampling_lambda=8, mode="swap")
    s.criterion = crit_mod
    real_model = small_net(3).train()
    other = small_net(4).train()
    ema = oracle.EMA(other, beta=0.9999, update_every=10, include_online_model=False)
    torch.manual_seed(302)
    out["flow_fake_swap_loss"] = s.training_swap_step_for_one_model("a", x, real_model, ema).detach()
    out["flow_fake_swap_diff"] = s.logged["swap_difference/a"]
    out["flow_fake_swap_ema_step"] = ema.step.clone()

    path = os.path.join(HERE, "ref_golden.npz")
    np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()})
    print("wrote", path, {k: tuple(np.shape(v)) for k, v in out.items()})


if __name__ == "__main__":
    main()
