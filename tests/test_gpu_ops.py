"""Per-op parity (-m gpu): every libd3fk kernel, called through its extern "C" entry point, against the
torch-CPU semantics of the same op on identical inputs.  Tolerances: fp32 engine 1e-5 (norm-relative),
bf16 engine 1e-2 on bf16-rounded inputs (fp32 accumulate, bf16 output rounding = 2^-9)."""
import math

import pytest
import torch

from denoising_diffusion_deep_fake_b200 import _lib
from gpu_harness import run_both, rel_err
import op_interpreter as I

pytestmark = pytest.mark.gpu
TOL = {_lib.F32: 1e-5, _lib.BF16: 1e-2}
DTYPES = [_lib.F32, _lib.BF16]


def _conv_case(dtype, B, Hi, Wi, c0, c1, up0, Cout, k, stride, pad, mode, relu=0, res=False, affine=False,
               stats=False, nchw=False, seed=0, splitk=False, bw=None, bias=False):
    g = torch.Generator().manual_seed(seed)
    ctot = c0 + c1
    if mode == 0:
        Ho = (Hi + 2 * pad - k) // stride + 1
        Wo = (Wi + 2 * pad - k) // stride + 1
    else:
        Ho, Wo = Hi * stride, Wi * stride
    hs, ws = (Hi >> up0), (Wi >> up0)
    t = {"src0": torch.randn(B, hs, ws, c0, generator=g), "w": torch.randn(Cout, k * k * ctot, generator=g) / math.sqrt(k * k * ctot)}
    if c1:
        t["src1"] = torch.randn(B, Hi, Wi, c1, generator=g)
    sc = dict(mode=mode, c0=c0, c1=c1, ld0=c0, ld1=c1, up0=up0, B=B, Hi=Hi, Wi=Wi, Ho=Ho, Wo=Wo, kh=k, kw=k,
              stride=stride, pad=pad, Cout=Cout, relu=relu)
    outs = []
    if nchw:
        t["out_nchw"] = torch.zeros(B, Cout, Ho, Wo)
        outs.append("out_nchw")
    else:
        t["out"] = torch.zeros(B, Ho, Wo, Cout)
        sc["ldo"] = Cout
        outs.append("out")
    if res:
        t["res"] = torch.randn(B, Ho, Wo, Cout, generator=g)
        sc["ldr"] = Cout
    if affine:
        t["scale"] = torch.rand(Cout, generator=g) + 0.5
        t["shift"] = torch.randn(Cout, generator=g)
    if bias:               # shift without scale: the segmentation head's bias
        t["shift"] = torch.randn(Cout, generator=g)
    if stats:
        t["stats"] = torch.zeros(2, Cout, dtype=torch.float64)
        outs.append("stats")
    if splitk:
        t["ws"] = torch.zeros(4 << 20)
        sc["ws_bytes"] = 16 << 20
    if bw is not None:     # fused BN-backward reduction (dgrad epilogue): bw = relu flag of the layer whose gradient this is
        t["bw_x"] = torch.randn(B, Ho, Wo, Cout, generator=g) * 1.5 + 0.3
        t["bw_act"] = torch.randn(B, Ho, Wo, Cout, generator=g).clamp_min(0)
        t["bw_mean"] = torch.randn(Cout, generator=g) * 0.3
        t["bw_invstd"] = torch.rand(Cout, generator=g) + 0.5
        t["stats"] = torch.zeros(2, Cout, dtype=torch.float64)
        sc.update(bw_ldx=Cout, bw_ldact=Cout, bw_relu=bw)
        outs.append("stats")
    return run_both(_lib.OP_CONV, dtype, t, sc, outs)


CONV_CASES = [
    # B, Hi, Wi, c0, c1, up0, Cout, k, stride, pad, mode, extras
    dict(B=2, Hi=16, Wi=16, c0=64, c1=0, up0=0, Cout=64, k=3, stride=1, pad=1, mode=0, stats=True),
    dict(B=3, Hi=8, Wi=8, c0=128, c1=0, up0=0, Cout=128, k=3, stride=1, pad=1, mode=0, relu=1, res=True, affine=True),
    dict(B=2, Hi=16, Wi=16, c0=64, c1=0, up0=0, Cout=128, k=3, stride=2, pad=1, mode=0, stats=True),
    dict(B=2, Hi=16, Wi=16, c0=64, c1=0, up0=0, Cout=128, k=1, stride=2, pad=0, mode=0),
    dict(B=2, Hi=64, Wi=64, c0=8, c1=0, up0=0, Cout=64, k=7, stride=2, pad=3, mode=0, stats=True),
    dict(B=2, Hi=8, Wi=8, c0=256, c1=128, up0=1, Cout=128, k=3, stride=1, pad=1, mode=0, stats=True),
    dict(B=1, Hi=32, Wi=32, c0=64, c1=64, up0=1, Cout=32, k=3, stride=1, pad=1, mode=0, relu=1, affine=True),
    dict(B=1, Hi=64, Wi=64, c0=32, c1=0, up0=1, Cout=16, k=3, stride=1, pad=1, mode=0, stats=True),
    dict(B=2, Hi=32, Wi=32, c0=16, c1=0, up0=0, Cout=3, k=3, stride=1, pad=1, mode=0, nchw=True),
    dict(B=5, Hi=2, Wi=2, c0=512, c1=0, up0=0, Cout=512, k=3, stride=1, pad=1, mode=0),
    dict(B=1, Hi=96, Wi=32, c0=16, c1=0, up0=0, Cout=16, k=3, stride=1, pad=1, mode=0, stats=True),   # ragged M tail
    # split-K (deep layers: few output tiles, long reduction)
    dict(B=16, Hi=4, Wi=4, c0=256, c1=0, up0=0, Cout=256, k=3, stride=1, pad=1, mode=0, stats=True, splitk=True),
    dict(B=9, Hi=2, Wi=2, c0=512, c1=0, up0=0, Cout=512, k=3, stride=1, pad=1, mode=0, relu=1, res=True, affine=True, splitk=True),
    dict(B=4, Hi=4, Wi=4, c0=512, c1=256, up0=1, Cout=256, k=3, stride=1, pad=1, mode=0, stats=True, splitk=True),
    dict(B=8, Hi=2, Wi=2, c0=512, c1=0, up0=0, Cout=256, k=3, stride=2, pad=1, mode=1, res=True, splitk=True),
    # many tiles per CTA (persistent tile loop)
    dict(B=16, Hi=64, Wi=64, c0=16, c1=0, up0=0, Cout=16, k=3, stride=1, pad=1, mode=0, stats=True),
    # dgrad (transposed gather)
    dict(B=2, Hi=16, Wi=16, c0=64, c1=0, up0=0, Cout=64, k=3, stride=1, pad=1, mode=1, res=True),
    dict(B=2, Hi=8, Wi=8, c0=128, c1=0, up0=0, Cout=64, k=3, stride=2, pad=1, mode=1),
    dict(B=2, Hi=8, Wi=8, c0=128, c1=0, up0=0, Cout=64, k=1, stride=2, pad=0, mode=1, res=True),
    dict(B=2, Hi=32, Wi=32, c0=8, c1=0, up0=0, Cout=16, k=3, stride=1, pad=1, mode=1),
    # slab path (3x3 s1, Cin in {16,32,64}, W in {16,32,64}): every swizzle width, forward and dgrad, all epilogues
    dict(B=3, Hi=32, Wi=32, c0=32, c1=0, up0=0, Cout=32, k=3, stride=1, pad=1, mode=0, stats=True),
    dict(B=3, Hi=32, Wi=32, c0=32, c1=0, up0=0, Cout=32, k=3, stride=1, pad=1, mode=1, res=True),
    dict(B=2, Hi=32, Wi=32, c0=32, c1=0, up0=0, Cout=64, k=3, stride=1, pad=1, mode=1),
    dict(B=2, Hi=64, Wi=64, c0=16, c1=0, up0=0, Cout=32, k=3, stride=1, pad=1, mode=1),
    dict(B=2, Hi=64, Wi=64, c0=16, c1=0, up0=0, Cout=16, k=3, stride=1, pad=1, mode=0, relu=1, res=True, affine=True),
    dict(B=5, Hi=16, Wi=16, c0=64, c1=0, up0=0, Cout=128, k=3, stride=1, pad=1, mode=1),
    dict(B=5, Hi=16, Wi=16, c0=64, c1=0, up0=0, Cout=64, k=3, stride=1, pad=1, mode=0, stats=True, seed=3),
    dict(B=40, Hi=16, Wi=16, c0=64, c1=0, up0=0, Cout=64, k=3, stride=1, pad=1, mode=0, stats=True, seed=4),   # > 1 tile per CTA
    dict(B=3, Hi=6, Wi=64, c0=16, c1=0, up0=0, Cout=16, k=3, stride=1, pad=1, mode=0, stats=True),            # S = 1 fallback
    dict(B=1, Hi=24, Wi=32, c0=32, c1=0, up0=0, Cout=32, k=3, stride=1, pad=1, mode=0, stats=True),           # S = 2
    dict(B=2, Hi=64, Wi=64, c0=16, c1=0, up0=0, Cout=3, k=3, stride=1, pad=1, mode=0, nchw=True),
    dict(B=2, Hi=8, Wi=128, c0=16, c1=0, up0=0, Cout=16, k=3, stride=1, pad=1, mode=0, stats=True),            # W = 128: one row per sub-tile
    dict(B=1, Hi=12, Wi=256, c0=32, c1=0, up0=0, Cout=32, k=3, stride=1, pad=1, mode=1, res=True),             # two tiles across W
    dict(B=1, Hi=4, Wi=256, c0=16, c1=0, up0=0, Cout=3, k=3, stride=1, pad=1, mode=0, nchw=True),
    dict(B=2, Hi=32, Wi=32, c0=16, c1=0, up0=0, Cout=16, k=3, stride=1, pad=1, mode=1),                        # head dgrad (dy padded to 16)
    dict(B=3, Hi=32, Wi=32, c0=128, c1=0, up0=0, Cout=32, k=3, stride=1, pad=1, mode=0, stats=True),            # two 64-channel chunks
    dict(B=2, Hi=16, Wi=16, c0=128, c1=0, up0=0, Cout=64, k=3, stride=1, pad=1, mode=0, relu=1, affine=True),
    dict(B=20, Hi=32, Wi=32, c0=128, c1=0, up0=0, Cout=32, k=3, stride=1, pad=1, mode=0, stats=True, seed=5),   # several tiles per CTA
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv(dtype, case):
    r = _conv_case(dtype, **case)
    for name, (g, c) in r.items():
        tol = TOL[dtype] if name != "stats" else (1e-5 if dtype == _lib.F32 else 2e-2)
        assert rel_err(g, c) < tol, (name, rel_err(g, c))


WGRAD_CASES = [
    dict(B=2, Hi=16, Wi=16, c0=64, c1=0, up0=0, Cout=64, k=3, stride=1, pad=1),
    dict(B=2, Hi=16, Wi=16, c0=64, c1=0, up0=0, Cout=128, k=3, stride=2, pad=1),
    dict(B=2, Hi=16, Wi=16, c0=64, c1=0, up0=0, Cout=128, k=1, stride=2, pad=0),
    dict(B=2, Hi=64, Wi=64, c0=8, c1=0, up0=0, Cout=64, k=7, stride=2, pad=3, cin_real=3),
    dict(B=2, Hi=8, Wi=8, c0=256, c1=128, up0=1, Cout=128, k=3, stride=1, pad=1),
    dict(B=1, Hi=64, Wi=64, c0=32, c1=0, up0=1, Cout=16, k=3, stride=1, pad=1),
    dict(B=2, Hi=32, Wi=32, c0=16, c1=0, up0=0, Cout=8, k=3, stride=1, pad=1, cout_real=3),
    dict(B=7, Hi=2, Wi=2, c0=512, c1=0, up0=0, Cout=512, k=3, stride=1, pad=1),
    dict(B=3, Hi=24, Wi=8, c0=16, c1=0, up0=0, Cout=16, k=3, stride=1, pad=1),
    # slab weight-gradient path (3x3 s1, Cin in {16,32,64}, Cout in {16,32}, >= 75 776 pixels)
    dict(B=19, Hi=64, Wi=64, c0=16, c1=0, up0=0, Cout=16, k=3, stride=1, pad=1),
    dict(B=19, Hi=64, Wi=64, c0=16, c1=0, up0=0, Cout=16, k=3, stride=1, pad=1, cout_real=3),
    dict(B=76, Hi=32, Wi=32, c0=32, c1=0, up0=0, Cout=32, k=3, stride=1, pad=1),
    dict(B=76, Hi=32, Wi=32, c0=64, c1=0, up0=0, Cout=32, k=3, stride=1, pad=1),
    dict(B=5, Hi=128, Wi=128, c0=16, c1=0, up0=0, Cout=32, k=3, stride=1, pad=1),
    dict(B=300, Hi=16, Wi=16, c0=64, c1=0, up0=0, Cout=16, k=3, stride=1, pad=1),
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", WGRAD_CASES)
def test_wgrad(dtype, case):
    c = dict(case)
    B, Hi, Wi, c0, c1, up0, Cout, k, stride, pad = (c[x] for x in ("B", "Hi", "Wi", "c0", "c1", "up0", "Cout", "k", "stride", "pad"))
    cin_real = c.get("cin_real", c0 + c1)
    cout_real = c.get("cout_real", Cout)
    g = torch.Generator().manual_seed(1)
    Ho = (Hi + 2 * pad - k) // stride + 1
    Wo = (Wi + 2 * pad - k) // stride + 1
    t = {"src0": torch.randn(B, Hi >> up0, Wi >> up0, c0, generator=g), "dy": torch.randn(B, Ho, Wo, Cout, generator=g),
         "dw": torch.zeros(cout_real, cin_real, k, k)}
    if c1:
        t["src1"] = torch.randn(B, Hi, Wi, c1, generator=g)
    sc = dict(c0=c0, c1=c1, ld0=c0, ld1=c1, up0=up0, B=B, Hi=Hi, Wi=Wi, Ho=Ho, Wo=Wo, kh=k, kw=k, stride=stride, pad=pad,
              ldy=Cout, Cout=Cout, cin_real=cin_real, cout_real=cout_real)
    r = run_both(_lib.OP_WGRAD, dtype, t, sc, ["dw"])
    gpu, cpu = r["dw"]
    assert rel_err(gpu, cpu) < (2e-5 if dtype == _lib.F32 else 1e-2), rel_err(gpu, cpu)


@pytest.mark.parametrize("dtype", DTYPES)
def test_pack_and_layout(dtype):
    g = torch.Generator().manual_seed(2)
    t = {"w": torch.randn(64, 3, 7, 7, generator=g), "w_fwd": torch.zeros(64, 49 * 8), "w_dgrad": torch.zeros(3, 49 * 64)}
    r = run_both(_lib.OP_PACK, dtype, t, dict(Cout=64, Cin=3, kh=7, kw=7, cin_pad=8, cout_pad=64), ["w_fwd", "w_dgrad"])
    for n, (a, b) in r.items():
        assert rel_err(a, b) < (1e-7 if dtype == _lib.F32 else 4e-3), n
    t = {"src": torch.randn(3, 3, 32, 64, generator=g), "dst": torch.ones(3, 32, 64, 8)}
    r = run_both(_lib.OP_NCHW2NHWC, dtype, t, dict(B=3, C=3, H=32, W=64, cpad=8), ["dst"])
    assert rel_err(*r["dst"]) < (1e-7 if dtype == _lib.F32 else 4e-3)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("C,count,relu,res", [(64, 2 * 16 * 16, 1, True), (16, 3 * 64 * 64, 1, False), (512, 8, 0, False),
                                               (256, 1000, 1, False), (32, 50000, 1, True)])
def test_bn_train_fwd_bwd(dtype, C, count, relu, res):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(count, C, generator=g) * 2 + 0.5
    xr = x if dtype == _lib.F32 else x.bfloat16().float()
    stats = torch.stack([xr.double().sum(0), (xr.double() ** 2).sum(0)])
    t = {"x": x, "y": torch.zeros(count, C), "stats": stats, "gamma": torch.rand(C, generator=g) + 0.5,
         "beta": torch.randn(C, generator=g), "running_mean": torch.zeros(C), "running_var": torch.ones(C),
         "num_batches_tracked": torch.zeros(1, dtype=torch.int64), "scale": torch.zeros(C), "shift": torch.zeros(C),
         "mean": torch.zeros(C), "invstd": torch.zeros(C)}
    sc = dict(C=C, relu=relu, count=count, ldx=C, ldy=C, eps=1e-5, momentum=0.1)
    if res:
        t["res"] = torch.randn(count, C, generator=g)
        sc["ldr"] = C
    r = run_both(_lib.OP_BN_FINALIZE, dtype, t, sc, ["scale", "shift", "mean", "invstd", "running_mean", "running_var", "num_batches_tracked"])
    for n, (a, b) in r.items():
        assert rel_err(a, b) < 1e-6, n
    # fused finalize + apply (what the plans run): statistics in, activation + saved mean/invstd + running stats out
    r = run_both(_lib.OP_BN_APPLY, dtype, t, sc, ["y", "mean", "invstd", "running_mean", "running_var", "num_batches_tracked"])
    assert rel_err(*r["y"]) < (1e-6 if dtype == _lib.F32 else 4e-3)
    for n in ("mean", "invstd", "running_mean", "running_var", "num_batches_tracked"):
        assert rel_err(*r[n]) < 1e-6, n
    # stand-alone apply with explicit scale/shift (eval-style)
    t2 = dict(t)
    t2.pop("stats")
    t2["scale"], t2["shift"] = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    r2 = run_both(_lib.OP_BN_APPLY, dtype, t2, sc, ["y"])
    assert rel_err(*r2["y"]) < (1e-6 if dtype == _lib.F32 else 4e-3)
    t["mean"], t["invstd"] = r["mean"][1], r["invstd"][1]
    # backward
    t["act"] = r["y"][1]
    t["dy"] = torch.randn(count, C, generator=g)
    t["bstats"] = torch.zeros(2, C, dtype=torch.float64)
    t["dgamma"], t["dbeta"], t["coef"] = torch.zeros(C), torch.zeros(C), torch.zeros(3, C)
    t["dx"], t["dres"] = torch.zeros(count, C), torch.zeros(count, C)
    sc.update(lddy=C, ldact=C, lddx=C, lddres=C)
    t.pop("res", None), t.pop("y"), t.pop("stats")
    sc.pop("ldr", None)
    r = run_both(_lib.OP_BN_BWD_REDUCE, dtype, t, sc, ["bstats"])
    assert rel_err(*r["bstats"]) < (1e-6 if dtype == _lib.F32 else 1e-4)
    t["bstats"] = r["bstats"][1].double()
    r = run_both(_lib.OP_BN_BWD_FINALIZE, dtype, t, sc, ["dgamma", "dbeta", "coef"])
    for n, (a, b) in r.items():
        assert rel_err(a, b) < 1e-6, n
    t["dgamma"], t["dbeta"] = torch.zeros(C), torch.zeros(C)
    r = run_both(_lib.OP_BN_BWD_APPLY, dtype, t, sc, ["dx", "dres", "dgamma", "dbeta"])
    for n, (a, b) in r.items():
        assert rel_err(a, b) < (1e-5 if dtype == _lib.F32 else 6e-3), n
    # the fused form the plans run: reduce + grid barrier + apply in one launch (zeroed sums and barrier counter in)
    t["bstats"] = torch.zeros(2, C, dtype=torch.float64)
    t["barrier"] = torch.zeros(2, dtype=torch.int32)
    t["dgamma"], t["dbeta"] = torch.zeros(C), torch.zeros(C)
    t["dx"], t["dres"] = torch.zeros(count, C), torch.zeros(count, C)
    r = run_both(_lib.OP_BN_BWD, dtype, t, sc, ["dx", "dres", "dgamma", "dbeta", "bstats"])
    for n, (a, b) in r.items():
        assert rel_err(a, b) < (1e-5 if dtype == _lib.F32 else 6e-3), n


@pytest.mark.parametrize("dtype", DTYPES)
def test_bn_fold(dtype):
    C = 128
    g = torch.Generator().manual_seed(4)
    t = {"gamma": torch.rand(C, generator=g) + 0.5, "beta": torch.randn(C, generator=g), "running_mean": torch.randn(C, generator=g),
         "running_var": torch.rand(C, generator=g) + 0.1, "scale": torch.zeros(C), "shift": torch.zeros(C)}
    r = run_both(_lib.OP_BN_FOLD, dtype, t, dict(C=C, eps=1e-5), ["scale", "shift"])
    for n, (a, b) in r.items():
        assert rel_err(a, b) < 1e-6, n


@pytest.mark.parametrize("dtype", DTYPES)
def test_pools(dtype):
    g = torch.Generator().manual_seed(5)
    B, H, W, C = 2, 32, 32, 64
    x = torch.relu(torch.randn(B, H, W, C, generator=g))         # many exact ties at 0, like post-ReLU features
    t = {"x": x, "y": torch.zeros(B, H // 2, W // 2, C), "idx": torch.zeros(B, H // 2, W // 2, C, dtype=torch.uint8)}
    sc = dict(B=B, H=H, W=W, C=C, ldx=C, ldy=C)
    r = run_both(_lib.OP_MAXPOOL_FWD, dtype, t, sc, ["y", "idx"])
    assert rel_err(*r["y"]) == 0.0
    assert torch.equal(r["idx"][0], r["idx"][1])
    t = {"dy": torch.randn(B, H // 2, W // 2, C, generator=g), "idx": r["idx"][1].to(torch.uint8), "dx": torch.randn(B, H, W, C, generator=g)}
    for acc in (0, 1):
        r2 = run_both(_lib.OP_MAXPOOL_BWD, dtype, t, dict(B=B, H=H, W=W, C=C, lddy=C, lddx=C, accumulate=acc), ["dx"])
        assert rel_err(*r2["dx"]) < (1e-6 if dtype == _lib.F32 else 5e-3)
    t = {"dy": torch.randn(B, H, W, C, generator=g), "dx": torch.zeros(B, H // 2, W // 2, C)}
    r3 = run_both(_lib.OP_SUMPOOL2, dtype, t, dict(B=B, H=H // 2, W=W // 2, C=C, lddy=C, lddx=C, accumulate=0), ["dx"])
    assert rel_err(*r3["dx"]) < (1e-6 if dtype == _lib.F32 else 5e-3)
    t = {"x": torch.randn(B * H * W, 8, generator=g), "out": torch.zeros(3)}
    r4 = run_both(_lib.OP_CHANSUM, dtype, t, dict(C=3, ld=8, count=B * H * W), ["out"])
    assert rel_err(*r4["out"]) < 1e-4


CONV_BN_CASES = [
    # fused conv + BN: 128-wide tiles, one resident tile per CTA; PATH 2 (TMA), PATH 0 (stride 2, 1x1), PATH 1 (upsample + concat)
    dict(B=16, Hi=8, Wi=8, c0=128, c1=0, up0=0, Cout=128, k=3, stride=1, pad=1, relu=1, res=True),
    dict(B=40, Hi=8, Wi=8, c0=128, c1=0, up0=0, Cout=128, k=3, stride=1, pad=1, relu=1, res=False),
    dict(B=16, Hi=4, Wi=4, c0=256, c1=0, up0=0, Cout=256, k=3, stride=1, pad=1, relu=1, res=True),      # split-K cluster
    dict(B=32, Hi=2, Wi=2, c0=512, c1=0, up0=0, Cout=512, k=3, stride=1, pad=1, relu=0, res=False),     # split-K cluster
    dict(B=8, Hi=16, Wi=16, c0=64, c1=0, up0=0, Cout=128, k=3, stride=2, pad=1, relu=1, res=False),
    dict(B=8, Hi=16, Wi=16, c0=64, c1=0, up0=0, Cout=128, k=1, stride=2, pad=0, relu=0, res=False),
    dict(B=4, Hi=8, Wi=8, c0=256, c1=128, up0=1, Cout=128, k=3, stride=1, pad=1, relu=1, res=False),
    dict(B=2, Hi=16, Wi=16, c0=64, c1=0, up0=0, Cout=64, k=3, stride=1, pad=1, relu=1, res=False),       # not fusable: two kernels
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", CONV_BN_CASES)
def test_conv_bn_op(dtype, case):
    """D3FK_OP_CONV_BN (conv -> train-mode BN -> +res -> ReLU, fused behind a grid barrier when the layer qualifies) vs
    conv then bn_apply on the CPU: raw output, activation, saved mean / invstd and running statistics."""
    import op_interpreter as I
    _lib.init(0)
    c = dict(case)
    B, Hi, Wi, c0, c1, up0, Cout, k, stride, pad = (c[x] for x in ("B", "Hi", "Wi", "c0", "c1", "up0", "Cout", "k", "stride", "pad"))
    g = torch.Generator().manual_seed(7)
    ctot = c0 + c1
    Ho = (Hi + 2 * pad - k) // stride + 1
    Wo = (Wi + 2 * pad - k) // stride + 1
    tdt = torch.float32 if dtype == _lib.F32 else torch.bfloat16
    host = {"src0": torch.randn(B, Hi >> up0, Wi >> up0, c0, generator=g),
            "w": torch.randn(Cout, k * k * ctot, generator=g) / math.sqrt(k * k * ctot),
            "raw": torch.zeros(B, Ho, Wo, Cout), "act": torch.zeros(B, Ho, Wo, Cout)}
    if c1:
        host["src1"] = torch.randn(B, Hi, Wi, c1, generator=g)
    if c["res"]:
        host["res"] = torch.randn(B, Ho, Wo, Cout, generator=g)
    f32 = {"stats": torch.zeros(2, Cout, dtype=torch.float64), "gamma": torch.rand(Cout, generator=g) + 0.5,
           "beta": torch.randn(Cout, generator=g), "running_mean": torch.zeros(Cout), "running_var": torch.ones(Cout),
           "nbt": torch.zeros(1, dtype=torch.int64), "mean": torch.zeros(Cout), "invstd": torch.zeros(Cout),
           "scale": torch.zeros(Cout), "shift": torch.zeros(Cout), "barrier": torch.zeros(2, dtype=torch.int32)}

    def build(t, s, dt):
        conv = dict(dtype=dt, mode=0, src0=t["src0"].data_ptr(), c0=c0, c1=c1, ld0=c0, ld1=c1, up0=up0, B=B, Hi=Hi, Wi=Wi, Ho=Ho,
                    Wo=Wo, kh=k, kw=k, stride=stride, pad=pad, w=t["w"].data_ptr(), Cout=Cout, out=t["raw"].data_ptr(), ldo=Cout,
                    stats=s["stats"].data_ptr())
        if c1:
            conv["src1"] = t["src1"].data_ptr()
        bn = dict(dtype=dt, C=Cout, relu=c["relu"], count=B * Ho * Wo, x=t["raw"].data_ptr(), ldx=Cout, y=t["act"].data_ptr(),
                  ldy=Cout, stats=s["stats"].data_ptr(), gamma=s["gamma"].data_ptr(), beta=s["beta"].data_ptr(),
                  running_mean=s["running_mean"].data_ptr(), running_var=s["running_var"].data_ptr(),
                  num_batches_tracked=s["nbt"].data_ptr(), eps=1e-5, momentum=0.1, scale=s["scale"].data_ptr(),
                  shift=s["shift"].data_ptr(), mean=s["mean"].data_ptr(), invstd=s["invstd"].data_ptr())
        if c["res"]:
            bn.update(res=t["res"].data_ptr(), ldr=Cout)
        return _lib.make_op(_lib.OP_CONV_BN, conv=conv, bn=bn, barrier=s["barrier"].data_ptr())

    gt = {n: v.to("cuda:0").to(tdt) for n, v in host.items()}
    gs = {n: v.to("cuda:0") for n, v in f32.items()}
    ct = {n: gt[n].float().cpu().contiguous() for n in host}
    cs = {n: v.clone() for n, v in f32.items()}
    _lib.run_single(build(gt, gs, dtype), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert _lib.load().d3fk_device_error_flag() == 0, "kernel watchdog tripped"
    I.run_ops([build(ct, cs, _lib.F32)])
    tol = 1e-5 if dtype == _lib.F32 else 1e-2
    assert rel_err(gt["raw"].float().cpu(), ct["raw"]) < tol
    assert rel_err(gt["act"].float().cpu(), ct["act"]) < (1e-5 if dtype == _lib.F32 else 2e-2)
    for n in ("mean", "invstd", "running_mean", "running_var"):
        assert rel_err(gs[n].cpu(), cs[n]) < (1e-5 if dtype == _lib.F32 else 5e-3), n
    assert int(gs["nbt"].item()) == 1


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("c0,c1", [(64, 64), (32, 0), (8, 24)])
def test_upcat(dtype, c0, c1):
    """Materialised nearest-2x upsample + channel concat (smp DecoderBlock) — exact copy semantics."""
    g = torch.Generator().manual_seed(9)
    B, H, W = 3, 16, 24
    t = {"src0": torch.randn(B, H // 2, W // 2, c0, generator=g), "out": torch.zeros(B, H, W, c0 + c1)}
    sc = dict(B=B, H=H, W=W, c0=c0, c1=c1, ld0=c0, ld1=c1, ldo=c0 + c1)
    if c1:
        t["src1"] = torch.randn(B, H, W, c1, generator=g)
    r = run_both(_lib.OP_UPCAT, dtype, t, sc, ["out"])
    assert torch.equal(r["out"][0], r["out"][1])


BW_CASES = [
    # BN-backward reduction fused into the dgrad epilogue: generic tiles, split-K cluster, stride 2, slab (N = 64 / 32 / 16)
    dict(B=16, Hi=8, Wi=8, c0=128, c1=0, up0=0, Cout=128, k=3, stride=1, pad=1, mode=1, res=True, bw=1),
    dict(B=16, Hi=4, Wi=4, c0=256, c1=0, up0=0, Cout=256, k=3, stride=1, pad=1, mode=1, bw=1),
    dict(B=16, Hi=4, Wi=4, c0=256, c1=0, up0=0, Cout=256, k=3, stride=1, pad=1, mode=1, res=True, bw=0),
    dict(B=4, Hi=8, Wi=8, c0=128, c1=0, up0=0, Cout=64, k=3, stride=2, pad=1, mode=1, bw=1),
    dict(B=6, Hi=16, Wi=16, c0=64, c1=0, up0=0, Cout=64, k=3, stride=1, pad=1, mode=1, res=True, bw=1),
    dict(B=3, Hi=32, Wi=32, c0=32, c1=0, up0=0, Cout=32, k=3, stride=1, pad=1, mode=1, bw=1),
    dict(B=2, Hi=64, Wi=64, c0=16, c1=0, up0=0, Cout=16, k=3, stride=1, pad=1, mode=1, bw=1),
    dict(B=2, Hi=64, Wi=64, c0=16, c1=0, up0=0, Cout=32, k=3, stride=1, pad=1, mode=1, bw=0),
]


@pytest.mark.parametrize("case", BW_CASES)
def test_dgrad_with_fused_bn_backward_reduction(case):
    """d3fk_conv_params.bw_*: the dgrad epilogue also produces sum(g') and sum(g' * xhat) of the BatchNorm it feeds (bf16 engine)."""
    r = _conv_case(_lib.BF16, **case)
    assert rel_err(*r["out"]) < 1e-2
    g, c = r["stats"]
    assert rel_err(g, c) < 2e-2, rel_err(g, c)


WGRAD_GROUP_CASES = [
    # (count, B, H, C): the encoder stages' shapes at reduced batch, plus regimes that exercise every split / cluster choice:
    # many tiles (no split), few tiles (cluster split over the pixels), a single problem, the maximum group size
    (6, 4, 16, 64), (7, 8, 8, 128), (11, 16, 4, 256), (5, 32, 2, 512), (1, 16, 4, 256), (12, 2, 8, 64), (2, 64, 4, 128),
]


@pytest.mark.parametrize("variant", ["halves_accumulate", "misaligned_dw", "narrow_out_misaligned"])
def test_wgrad_slab_epilogue_paths(variant):
    """The slab weight gradient's epilogue stages the CTA's result in dW order and adds it with coalesced reductions
    (16-byte `red.global.add.v4.f32` when dW rows are 16-byte aligned, scalar otherwise).  Covered here: the two 64-channel
    halves of a 128-channel input accumulating into one NON-ZERO dW (plan.py `_wgrad_ops`), and dW at a 4-byte offset."""
    _lib.init(0)
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11)
    if variant == "halves_accumulate":
        B, H, W, cin, C, Cout, cout_real, off = 76, 32, 32, 128, 64, 32, 32, 0
    elif variant == "misaligned_dw":
        B, H, W, cin, C, Cout, cout_real, off = 76, 32, 32, 32, 32, 32, 32, 1
    else:
        B, H, W, cin, C, Cout, cout_real, off = 19, 64, 64, 16, 16, 16, 3, 3
    x = torch.randn(B, H, W, cin, generator=g).bfloat16().to(dev)
    dy = torch.randn(B, H, W, Cout, generator=g).bfloat16().to(dev)
    n = cout_real * cin * 9
    arena = torch.ones(n + 8, device=dev)
    dw = arena[off:off + n]
    s = torch.cuda.current_stream().cuda_stream
    for half in range(cin // C):
        op = _lib.make_op(_lib.OP_WGRAD, dtype=_lib.BF16, src0=x.data_ptr() + half * C * 2, c0=C, c1=0, ld0=cin, ld1=0, up0=0,
                          B=B, Hi=H, Wi=W, Ho=H, Wo=W, kh=3, kw=3, stride=1, pad=1, dy=dy.data_ptr(), ldy=Cout, Cout=Cout,
                          cin_real=cin, cout_real=cout_real, dw=dw.data_ptr() + half * C * 9 * 4)
        _lib.run_single(op, s)
    torch.cuda.synchronize()
    assert _lib.load().d3fk_device_error_flag() == 0, "kernel watchdog tripped"
    ref = torch.nn.grad.conv2d_weight(x.float().cpu().permute(0, 3, 1, 2).double(), (Cout, cin, 3, 3),
                                      dy.float().cpu().permute(0, 3, 1, 2).double(), stride=1, padding=1)[:cout_real]
    got = dw.cpu().double().view(cout_real, cin, 3, 3) - 1.0        # dW started at one: the launch accumulates
    assert rel_err(got, ref) < 1e-4, rel_err(got, ref)
    assert float(arena[:off].sub(1).abs().sum() + arena[off + n:].sub(1).abs().sum()) == 0.0   # nothing outside dW touched


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("count,B,H,C", WGRAD_GROUP_CASES)
def test_wgrad_group(dtype, count, B, H, C):
    """d3fk_wgrad_group: `count` identically shaped 3x3 / stride-1 weight gradients in one launch, against the CPU interpreter
    (per problem) and against the single-problem entry point d3fk_wgrad on the same operands."""
    _lib.init(0)
    dev = "cuda:0"
    tdt = torch.float32 if dtype == _lib.F32 else torch.bfloat16
    g = torch.Generator().manual_seed(5)
    xs = [torch.randn(B, H, H, C, generator=g).to(dev).to(tdt) for _ in range(count)]
    dys = [torch.randn(B, H, H, C, generator=g).to(dev).to(tdt) for _ in range(count)]
    dws = [torch.zeros(C, C, 3, 3, device=dev) for _ in range(count)]
    base = dict(dtype=dtype, c0=C, c1=0, ld0=C, ld1=0, up0=0, B=B, Hi=H, Wi=H, Ho=H, Wo=H, kh=3, kw=3, stride=1, pad=1,
                ldy=C, Cout=C, cin_real=C, cout_real=C)
    op = _lib.make_op(_lib.OP_WGRAD_GROUP, base=base, count=count, src0=[t.data_ptr() for t in xs],
                      dy=[t.data_ptr() for t in dys], dw=[t.data_ptr() for t in dws])
    stream = torch.cuda.current_stream().cuda_stream
    _lib.run_single(op, stream)
    torch.cuda.synchronize()
    assert _lib.load().d3fk_device_error_flag() == 0, "kernel watchdog tripped"
    tol = 2e-5 if dtype == _lib.F32 else 1e-2
    for i in range(count):
        x_c, dy_c = xs[i].float().cpu().contiguous(), dys[i].float().cpu().contiguous()
        dw_c = torch.zeros(C, C, 3, 3)
        cpu_base = dict(base, dtype=_lib.F32)
        I.run_ops([_lib.make_op(_lib.OP_WGRAD, src0=x_c.data_ptr(), dy=dy_c.data_ptr(), dw=dw_c.data_ptr(), **cpu_base)])
        assert rel_err(dws[i].cpu(), dw_c) < tol, (i, rel_err(dws[i].cpu(), dw_c))
        one = torch.zeros(C, C, 3, 3, device=dev)
        _lib.run_single(_lib.make_op(_lib.OP_WGRAD, src0=xs[i].data_ptr(), dy=dys[i].data_ptr(), dw=one.data_ptr(), **base), stream)
        torch.cuda.synchronize()
        assert rel_err(dws[i].cpu(), one.cpu()) < (1e-6 if dtype == _lib.F32 else 1e-5)


HEAD_CASES = [
    # the segmentation head (16 -> 3, bias, fp32 NCHW out): CUDA-core kernel head_conv.cu for every tile width (W % 128 / 64 /
    # 32), one output channel, a single image, and shapes that must fall back to the tensor-core path (H % 32 != 0)
    dict(B=2, Hi=64, Wi=64, Cout=3), dict(B=1, Hi=128, Wi=128, Cout=3), dict(B=3, Hi=32, Wi=96, Cout=3),
    dict(B=2, Hi=32, Wi=32, Cout=1), dict(B=1, Hi=64, Wi=256, Cout=2), dict(B=2, Hi=16, Wi=64, Cout=3),
]


@pytest.mark.parametrize("case", HEAD_CASES)
@pytest.mark.parametrize("bias", [False, True])
def test_head_conv(case, bias):
    r = _conv_case(_lib.BF16, c0=16, c1=0, up0=0, k=3, stride=1, pad=1, mode=0, nchw=True, bias=bias, seed=3, **case)
    g, c = r["out_nchw"]
    assert rel_err(g, c) < 2e-3, rel_err(g, c)        # bf16 inputs / weights, fp32 accumulation on both sides
    # edges (zero padding) specifically
    assert rel_err(g[..., 0, :], c[..., 0, :]) < 2e-3 and rel_err(g[..., :, -1], c[..., :, -1]) < 2e-3


# ---- the cp.async (LDGSTS) operand path has no fence.proxy.async between the full-barrier wait and tcgen05.mma
# (csrc/tc_common.cuh, "NOTE on proxies"; the protocol of CUTLASS's sm100 cp.async main loop).  A generic->async proxy
# race would show up as run-to-run differences: ragged shapes through every gather path, 1000 launches in total, each
# compared bit for bit with the first.
CP_ASYNC_STRESS = [
    # conv PATH 0 (linear gather): strided forward, 1x1/2, the 7x7 stem, a ragged M tail
    ("conv", dict(B=3, Hi=16, Wi=16, c0=64, c1=0, up0=0, Cout=128, k=3, stride=2, pad=1, mode=0)),
    ("conv", dict(B=5, Hi=8, Wi=8, c0=128, c1=0, up0=0, Cout=256, k=1, stride=2, pad=0, mode=0)),
    ("conv", dict(B=1, Hi=32, Wi=96, c0=8, c1=0, up0=0, Cout=64, k=7, stride=2, pad=3, mode=0)),
    ("conv", dict(B=7, Hi=6, Wi=10, c0=64, c1=0, up0=0, Cout=64, k=3, stride=1, pad=1, mode=0)),
    # conv PATH 1 (generic gather): upsample + concat, stride-2 transposed gather
    ("conv", dict(B=3, Hi=8, Wi=8, c0=256, c1=128, up0=1, Cout=128, k=3, stride=1, pad=1, mode=0)),
    ("conv", dict(B=2, Hi=6, Wi=10, c0=128, c1=0, up0=0, Cout=64, k=3, stride=2, pad=1, mode=1)),
    # generic weight gradient (per-thread cp.async gather of both operands)
    ("wgrad", dict(B=3, Hi=16, Wi=16, c0=64, c1=0, up0=0, Cout=128, k=3, stride=2, pad=1)),
    ("wgrad", dict(B=2, Hi=8, Wi=8, c0=256, c1=128, up0=1, Cout=128, k=3, stride=1, pad=1)),
    ("wgrad", dict(B=5, Hi=6, Wi=10, c0=64, c1=0, up0=0, Cout=64, k=3, stride=1, pad=1)),
    ("wgrad", dict(B=9, Hi=2, Wi=2, c0=512, c1=0, up0=0, Cout=512, k=3, stride=1, pad=1)),
]


def test_cp_async_path_is_race_free():
    _lib.init(0)
    dev = "cuda:0"
    stream = torch.cuda.current_stream().cuda_stream
    iters = 1000 // len(CP_ASYNC_STRESS)
    for kind, c in CP_ASYNC_STRESS:
        g = torch.Generator().manual_seed(7)
        B, Hi, Wi, c0, c1, up0, Cout, k, stride, pad = (c[x] for x in ("B", "Hi", "Wi", "c0", "c1", "up0", "Cout", "k", "stride", "pad"))
        ctot = c0 + c1
        src0 = torch.randn(B, Hi >> up0, Wi >> up0, c0, generator=g).to(dev).bfloat16()
        src1 = torch.randn(B, Hi, Wi, max(c1, 8), generator=g).to(dev).bfloat16()
        if kind == "conv":
            mode = c["mode"]
            Ho, Wo = ((Hi + 2 * pad - k) // stride + 1, (Wi + 2 * pad - k) // stride + 1) if mode == 0 else (Hi * stride, Wi * stride)
            w = (torch.randn(Cout, k * k * ctot, generator=g) / math.sqrt(k * k * ctot)).to(dev).bfloat16()
            out = torch.zeros(B, Ho, Wo, Cout, device=dev, dtype=torch.bfloat16)
            op = _lib.make_op(_lib.OP_CONV, dtype=_lib.BF16, mode=mode, src0=src0.data_ptr(), src1=src1.data_ptr() if c1 else None,
                              c0=c0, c1=c1, ld0=c0, ld1=c1, up0=up0, B=B, Hi=Hi, Wi=Wi, Ho=Ho, Wo=Wo, kh=k, kw=k, stride=stride,
                              pad=pad, w=w.data_ptr(), Cout=Cout, out=out.data_ptr(), ldo=Cout)
            exact = True
        else:
            Ho, Wo = (Hi + 2 * pad - k) // stride + 1, (Wi + 2 * pad - k) // stride + 1
            dy = torch.randn(B, Ho, Wo, Cout, generator=g).to(dev).bfloat16()
            out = torch.zeros(Cout, ctot, k, k, device=dev)
            op = _lib.make_op(_lib.OP_WGRAD, dtype=_lib.BF16, src0=src0.data_ptr(), src1=src1.data_ptr() if c1 else None, c0=c0, c1=c1,
                              ld0=c0, ld1=c1, up0=up0, B=B, Hi=Hi, Wi=Wi, Ho=Ho, Wo=Wo, kh=k, kw=k, stride=stride, pad=pad,
                              dy=dy.data_ptr(), dw=out.data_ptr(), ldy=Cout, Cout=Cout, cin_real=ctot, cout_real=Cout)
            exact = False       # pixel splits beyond one cluster are summed with atomics: order-dependent last bits
        first = None
        bad = torch.zeros((), device=dev)
        for _ in range(iters):
            if kind == "wgrad":
                out.zero_()
            _lib.run_single(op, stream)
            if first is None:
                first = out.clone()
                assert torch.isfinite(first.float()).all() and first.float().abs().max() > 0
            elif exact:
                bad += (out != first).any()
            else:
                bad += ((out - first).norm() > 1e-5 * first.norm())
        assert bad.item() == 0, (kind, c, bad.item())
    assert _lib.load().d3fk_device_error_flag() == 0


@pytest.mark.parametrize("B,H,W", [(2, 64, 64), (3, 32, 32), (1, 128, 128), (1, 256, 256), (5, 64, 32)])
def test_stem_space_to_depth(B, H, W):
    """The 7x7 / stride-2 stem as a space-to-depth convolution (D3FK_OP_NCHW2S2D + D3FK_OP_PACK_STEM + conv mode 2: every
    (pixel, kh) one 128-byte window of the padded space-to-depth image, A operand by TMA over a tensor map with overlapping
    rows) against torch's conv2d on the same bf16-rounded operands, with the batch statistics of its epilogue."""
    import math
    import torch.nn.functional as F
    dev = "cuda:0"
    _lib.init(0)
    stream = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(B * 7 + H)
    x = torch.randn(B, 3, H, W, generator=g).to(dev)
    w = (torch.randn(64, 3, 7, 7, generator=g) / math.sqrt(147)).to(dev)
    Hs, Ws = H // 2, W // 2
    xs = torch.zeros(B, Hs, Ws + 3, 16, device=dev, dtype=torch.bfloat16)
    wp = torch.zeros(64, 256, device=dev, dtype=torch.bfloat16)
    out = torch.zeros(B, Hs, Ws, 64, device=dev, dtype=torch.bfloat16)
    st = torch.zeros(2, 64, dtype=torch.float64, device=dev)
    ops = [_lib.make_op(_lib.OP_NCHW2S2D, dtype=_lib.BF16, B=B, C=3, H=H, W=W, cpad=4, src=x.data_ptr(), dst=xs.data_ptr()),
           _lib.make_op(_lib.OP_PACK_STEM, dtype=_lib.BF16, Cout=64, Cin=3, kh=7, kw=7, cin_pad=4, cout_pad=64, w=w.data_ptr(),
                        w_fwd=wp.data_ptr()),
           _lib.make_op(_lib.OP_CONV, dtype=_lib.BF16, mode=2, src0=xs.data_ptr(), c0=64, c1=0, ld0=16, ld1=0, up0=0, B=B, Hi=Hs,
                        Wi=Ws, Ho=Hs, Wo=Ws, kh=4, kw=1, stride=1, pad=2, w=wp.data_ptr(), Cout=64, out=out.data_ptr(), ldo=64,
                        stats=st.data_ptr())]
    _lib.OpList(ops).run(stream)
    torch.cuda.synchronize()
    ref = F.conv2d(x.bfloat16().double(), w.bfloat16().double(), stride=2, padding=3).permute(0, 2, 3, 1)
    assert rel_err(out.float().cpu(), ref.float().cpu()) < 4e-3          # bf16 rounding of the stored output
    assert rel_err(st[0].cpu(), ref.sum((0, 1, 2)).cpu()) < 1e-5
    assert rel_err(st[1].cpu(), (ref * ref).sum((0, 1, 2)).cpu()) < 1e-5
    assert xs[:, :, :2].abs().max() == 0 and xs[:, :, -1].abs().max() == 0      # the pad pixels stay zero
    # weight gradient from the same windowed rows (d3fk_wgrad_params.mode 2) into the 7x7 OIHW master gradient
    dy = torch.randn(B, Hs, Ws, 64, generator=g).to(dev).bfloat16()
    dw = torch.zeros(64, 3, 7, 7, device=dev)
    _lib.run_single(_lib.make_op(_lib.OP_WGRAD, dtype=_lib.BF16, mode=2, src0=xs.data_ptr(), c0=64, c1=0, ld0=16, up0=0, B=B, Hi=Hs,
                                 Wi=Ws, Ho=Hs, Wo=Ws, kh=4, kw=1, stride=1, pad=2, dy=dy.data_ptr(), ldy=64, Cout=64, cin_real=3,
                                 cout_real=64, dw=dw.data_ptr()), stream)
    torch.cuda.synchronize()
    xd = x.bfloat16().double().cpu()
    dw_ref = torch.nn.grad.conv2d_weight(xd, (64, 3, 7, 7), dy.double().cpu().permute(0, 3, 1, 2), stride=2, padding=3)
    assert rel_err(dw.cpu(), dw_ref.float()) < 1e-4, rel_err(dw.cpu(), dw_ref.float())
    assert _lib.load().d3fk_device_error_flag() == 0
