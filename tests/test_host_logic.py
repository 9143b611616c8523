"""CPU suite: schedules, posterior coefficients, EMA semantics, CLI surface, sharding helpers and a
world_size-2 gloo run of the gradient bucket allreduce (the N>1 path of bench.py / train.py)."""
import math
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
import denoising_diffusion_deep_fake_b200 as d3
from denoising_diffusion_deep_fake_b200 import parallel
from denoising_diffusion_deep_fake_b200.train import EMA, cosine_lr


def test_posterior_coefficients_match_oracle():
    for eta in (0.0, 0.5, 1.0):
        grid = d3.noise_ratio_grid(17, r_start=1.0)
        gref = oracle.noise_ratio_grid(17).tolist()
        assert max(abs(a - b) for a, b in zip(grid, gref)) < 1e-12
        for i in range(17):
            a = d3.posterior_coeffs(grid[i], grid[i + 1], eta)
            b = oracle.sampler.posterior_coeffs(gref[i], gref[i + 1], eta)
            assert max(abs(u - v) for u, v in zip(a, b)) < 1e-12
    assert d3.posterior_coeffs(0.3, 0.0, 1.0) == (0.0, 1.0, 0.0)             # last step returns x0_hat


def test_cosine_lr_matches_torch_scheduler():
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=0.02)
    sch = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=100)
    for epoch in range(1, 50):
        opt.step()
        sch.step()
        assert abs(opt.param_groups[0]["lr"] - cosine_lr(0.02, epoch, 100)) < 1e-9


def test_product_ema_matches_oracle_ema():
    torch.manual_seed(0)
    a, b = torch.nn.Linear(6, 6), torch.nn.Linear(6, 6)
    b.load_state_dict(a.state_dict())
    ea = oracle.EMA(a, beta=0.999, update_every=1, include_online_model=False)
    eb = EMA(b, beta=0.999, update_every=1, include_online_model=False)
    for _ in range(130):
        with torch.no_grad():
            d = torch.randn(6, 6) * 0.1
            a.weight.add_(d), b.weight.add_(d)
        ea.update(), eb.update()
        assert torch.allclose(ea.ema_model.weight, eb.ema_model.weight, atol=1e-6)
    assert set(ea.state_dict()) == set(eb.state_dict())                       # ema_model.*, initted, step


def test_cli_surface():
    from click.testing import CliRunner
    from denoising_diffusion_deep_fake_b200.main import cli
    out = CliRunner().invoke(cli, ["--help"]).output
    for cmd in ("denoise", "train", "sample"):
        assert cmd in out
    out = CliRunner().invoke(cli, ["train", "--help"]).output
    for cmd in ("new", "resume", "modify"):
        assert cmd in out
    assert "--config_path" in CliRunner().invoke(cli, ["train", "new", "--help"]).output
    assert "--input_list" in CliRunner().invoke(cli, ["denoise", "--help"]).output


def test_shard_range_partitions():
    for total in (0, 1, 7, 64, 257):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1


def _dp_worker(rank, world, port, buckets, n):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        arena = torch.randn(n, generator=g)
        expect = sum(torch.randn(n, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)) / world
        for s, e in buckets:                                  # one allreduce per backward segment, as in training
            parallel.allreduce_bucket_(arena, s, e)
        assert torch.allclose(arena, expect, atol=1e-6)
        flat = torch.full((5,), float(rank))
        parallel.broadcast_flat_(flat, src=0)
        assert torch.equal(flat, torch.zeros(5))
        x = torch.arange(10.0).view(10, 1)
        mine = parallel.shard_batch(x, world, rank)
        allx = parallel.gather_shards(mine.contiguous(), world)
        assert torch.equal(allx, x)
    finally:
        dist.destroy_process_group()


def test_gradient_bucket_allreduce_gloo_world2():
    m = d3.Unet(precision="fp32")
    m._ensure_param_tables()
    buckets = m.grad_buckets()
    n = m._grad_numel
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_dp_worker, args=(2, port, buckets, n), nprocs=2, join=True)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm): one JSON line with the contract's keys,
    rank 0 only under a multi-rank launch — no GPU needed."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_img_per_s" and d["unit"] == "img/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"]
    # every other rank of a torchrun launch exits 0 without work and without output
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_ema_resume_keeps_schedule():
    """A checkpoint taken mid-run restores the host-side schedule mirrors: after load_state_dict the next update() must
    continue the decay warm-up where it stopped (not re-copy the online weights over the loaded EMA)."""
    torch.manual_seed(0)
    a, b = torch.nn.Linear(6, 6), torch.nn.Linear(6, 6)
    b.load_state_dict(a.state_dict())
    ea = oracle.EMA(a, beta=0.999, update_every=5, include_online_model=False)
    eb = EMA(b, beta=0.999, update_every=5, include_online_model=False)

    def drift():
        with torch.no_grad():
            d = torch.randn(6, 6) * 0.1
            a.weight.add_(d), b.weight.add_(d)

    for _ in range(137):
        drift(), ea.update(), eb.update()
    sd = {k: v.clone() for k, v in eb.state_dict().items()}
    resumed = EMA(b, beta=0.999, update_every=5, include_online_model=False)      # fresh object: mirrors at 0 / False
    resumed.load_state_dict(sd)
    assert resumed._step_host == 137 and resumed._initted_host is True
    for _ in range(40):
        drift(), ea.update(), resumed.update()
        assert torch.allclose(ea.ema_model.weight, resumed.ema_model.weight, atol=1e-6)
    assert int(resumed.step) == int(ea.step) == 177
    # as a submodule of a checkpointed parent (the way DeepFakeModule holds it)
    holder = torch.nn.Module()
    holder.ema_model_a = EMA(b, beta=0.999, update_every=5, include_online_model=False)
    holder.load_state_dict({"ema_model_a." + k: v for k, v in sd.items()})
    assert holder.ema_model_a._step_host == 137 and holder.ema_model_a._initted_host is True


def test_checkpoint_restores_optimizer_state_and_cosine_lr(tmp_path):
    """`train resume`: Adam moments / step counts come back from the checkpoint and the resumed epoch runs at
    cosine_lr(current_epoch), not at the base LR (Lightning's ckpt_path semantics)."""
    from denoising_diffusion_deep_fake_b200.main import load_checkpoint, restore_optimizers, save_checkpoint
    from denoising_diffusion_deep_fake_b200.train import DeepFakeModule
    hp = dict(encoder_name="resnet34", learning_rate=0.02, noise_exponential_sampling_lambda=3, max_epochs=50,
              cosine_scheduler_max_epoch=50, mode="denoise", adam_b1=0.5, adam_b2=0.999, batch_size=2, precision="fp32")
    m = DeepFakeModule(**hp)
    opts = m.configure_optimizers(fused=False)
    for o in opts:                                  # a fake step so that the state is non-trivial
        for p in o.param_groups[0]["params"][:3]:
            p.grad = torch.ones_like(p)
        o.step()
    m.current_epoch = 7
    path = str(tmp_path / "last.ckpt")
    save_checkpoint(path, m, epoch=7)
    m2, ckpt = load_checkpoint(path, DeepFakeModule)
    assert m2.current_epoch == 7 and len(ckpt["optimizer_states"]) == 2
    opts2 = m2.configure_optimizers(fused=False)
    restore_optimizers(m2, ckpt)
    for o, o2 in zip(opts, opts2):
        s, s2 = o.state_dict()["state"], o2.state_dict()["state"]
        assert set(s) == set(s2) and len(s) == 3
        for k in s:
            assert torch.equal(s[k]["exp_avg"], s2[k]["exp_avg"]) and float(s[k]["step"]) == float(s2[k]["step"]) == 1.0
        assert o2.param_groups[0]["lr"] == pytest.approx(cosine_lr(0.02, 7, 50))
        assert o2.param_groups[0]["betas"] == (0.5, 0.999)


def test_denoise_cli_needs_a_data_source(tmp_path):
    from click.testing import CliRunner
    from denoising_diffusion_deep_fake_b200.main import cli
    cfg = tmp_path / "c.yml"
    cfg.write_text("batch_size: 2\nlearning_rate: 0.02\nmax_epochs: 1\ncosine_scheduler_max_epoch: 1\nencoder_name: resnet34\n"
                   "noise_exponential_sampling_lambda: 5\n")
    res = CliRunner().invoke(cli, ["denoise", "--config", str(cfg)])
    assert res.exit_code != 0 and "--input_list" in res.output


def test_random_affine_maps_with_probability():
    from denoising_diffusion_deep_fake_b200.functional import random_affine_inverse_maps
    g = torch.Generator().manual_seed(0)
    m = random_affine_inverse_maps(4000, 32, 32, scale=(0.9, 1.1), p=0.7, generator=g)
    ident = torch.tensor([1.0, 0.0, 0.0, 0.0, 1.0, 0.0])
    frac = (m == ident).all(dim=1).float().mean().item()
    assert abs(frac - 0.3) < 0.03
    det = m[:, 0] * m[:, 4] - m[:, 1] * m[:, 3]                 # inverse map: 1 / scale^2
    assert (det > 1 / 1.1 ** 2 - 1e-4).all() and (det < 1 / 0.9 ** 2 + 1e-4).all()
