"""GPU probe: the FLAT slab-convolution variant (one slab per stage, tap = flat pixel-row offset of the UMMA descriptor)
against torch on the same bf16 operands, for every swizzle width, forward and dgrad, ragged sizes; then its speed against
the three-slab layout on the decoder-tail shapes.  Needs the debug build:
    D3FK_LIB=tools/libd3fk_dbg.so D3FK_SLAB_FLAT=1 D3FK_SLAB_BO=<0|1|2> python tools/probe_slab_flat.py"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from denoising_diffusion_deep_fake_b200 import _lib

dev = "cuda:0"
_lib.init(0)
stream = torch.cuda.current_stream().cuda_stream
tag = f"flat={os.environ.get('D3FK_SLAB_FLAT', 'default')} bo={os.environ.get('D3FK_SLAB_BO', 'default')}"


def run(B, H, W, C, Cout, mode, reps=0, res=False, stats=False):
    g = torch.Generator().manual_seed(B * 131 + H * 7 + W + C + Cout + mode)
    x = torch.randn(B, H, W, C, generator=g).to(dev).bfloat16()
    w = (torch.randn(Cout, 9 * C, generator=g) / math.sqrt(9 * C)).to(dev).bfloat16()
    out = torch.zeros(B, H, W, Cout, device=dev, dtype=torch.bfloat16)
    f = dict(dtype=_lib.BF16, mode=mode, src0=x.data_ptr(), c0=C, c1=0, ld0=C, ld1=0, up0=0, B=B, Hi=H, Wi=W, Ho=H, Wo=W,
             kh=3, kw=3, stride=1, pad=1, w=w.data_ptr(), Cout=Cout, out=out.data_ptr(), ldo=Cout)
    st = None
    if stats:
        st = torch.zeros(2, Cout, dtype=torch.float64, device=dev)
        f.update(stats=st.data_ptr())
    op = _lib.make_op(_lib.OP_CONV, **f)
    _lib.run_single(op, stream)
    torch.cuda.synchronize()
    xn = x.float().permute(0, 3, 1, 2)
    wk = w.float().view(Cout, 3, 3, C).permute(0, 3, 1, 2)         # [co][kh][kw][c] -> OIHW
    if mode == 0:
        ref = F.conv2d(xn, wk, padding=1)
    else:                                                           # transposed gather: out[ci] = sum A[h+1-kh, w+1-kw, co] w[ci][kh][kw][co]
        ref = F.conv2d(xn, wk.flip(2, 3), padding=1)
    ref = ref.permute(0, 2, 3, 1)
    err = ((out.float() - ref).norm() / ref.norm()).item()
    serr = None
    if stats:
        serr = ((st[0] - ref.double().sum((0, 1, 2))).norm() / ref.double().sum((0, 1, 2)).norm()).item()
    ms = None
    if reps:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ol = _lib.OpList([op] * reps)
        ol.run(stream); torch.cuda.synchronize()
        e0.record(); ol.run(stream); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
    return err, serr, ms


print("==", tag)
cases = [(2, 32, 32, 16, 16, 0), (2, 32, 32, 32, 32, 0), (2, 32, 32, 64, 64, 0), (2, 32, 32, 128, 32, 0),
         (2, 64, 64, 16, 16, 0), (2, 64, 64, 32, 16, 1), (1, 64, 64, 64, 64, 1), (3, 32, 32, 32, 64, 1),
         (1, 40, 96, 16, 16, 0), (1, 33, 48, 32, 32, 0), (1, 12, 256, 16, 16, 0), (2, 128, 128, 16, 16, 0)]
for c in cases:
    try:
        err, serr, _ = run(*c, stats=(c[5] == 0))
        print(f"B={c[0]} {c[1]}x{c[2]} C={c[3]} Cout={c[4]} mode={c[5]}: rel err {err:.3e}" + (f" stats err {serr:.3e}" if serr is not None else ""), flush=True)
    except Exception as ex:
        print("case", c, "FAILED:", str(ex)[:200], flush=True)
print("-- timing (20 back-to-back launches)")
for c in [(256, 64, 64, 32, 16, 0), (256, 64, 64, 16, 16, 0), (256, 32, 32, 128, 32, 0), (256, 32, 32, 32, 32, 0),
          (256, 64, 64, 16, 32, 1), (256, 64, 64, 16, 16, 1), (256, 32, 32, 32, 64, 1), (64, 128, 128, 32, 16, 0),
          (64, 128, 128, 16, 16, 0), (64, 64, 64, 128, 32, 0), (64, 32, 32, 64, 64, 0), (64, 32, 32, 64, 64, 1)]:
    try:
        err, serr, ms = run(*c, reps=20, stats=(c[5] == 0))
        print(f"B={c[0]} {c[1]}x{c[2]} C={c[3]} Cout={c[4]} mode={c[5]}: {ms * 1e3:7.1f} us  rel err {err:.3e}", flush=True)
    except Exception as ex:
        print("case", c, "FAILED:", str(ex)[:200], flush=True)
print("device error flag:", _lib.load().d3fk_device_error_flag())
