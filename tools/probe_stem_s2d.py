"""GPU probe: the space-to-depth stem (D3FK_OP_NCHW2S2D + D3FK_OP_PACK_STEM + conv mode 2) against torch's 7x7 / stride-2 conv,
and its time next to the gather form."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from denoising_diffusion_deep_fake_b200 import _lib

dev = "cuda:0"
_lib.init(0)
stream = torch.cuda.current_stream().cuda_stream

def run(B, H, W, reps=0):
    g = torch.Generator().manual_seed(B + H)
    x = torch.randn(B, 3, H, W, generator=g).to(dev)
    w = (torch.randn(64, 3, 7, 7, generator=g) / math.sqrt(147)).to(dev)
    Hs, Ws = H // 2, W // 2
    xs = torch.zeros(B, Hs, Ws + 3, 16, device=dev, dtype=torch.bfloat16)
    wp = torch.zeros(64, 256, device=dev, dtype=torch.bfloat16)
    out = torch.zeros(B, Hs, Ws, 64, device=dev, dtype=torch.bfloat16)
    st = torch.zeros(2, 64, dtype=torch.float64, device=dev)
    ops = [_lib.make_op(_lib.OP_NCHW2S2D, dtype=_lib.BF16, B=B, C=3, H=H, W=W, cpad=4, src=x.data_ptr(), dst=xs.data_ptr()),
           _lib.make_op(_lib.OP_PACK_STEM, dtype=_lib.BF16, Cout=64, Cin=3, kh=7, kw=7, cin_pad=4, cout_pad=64, w=w.data_ptr(), w_fwd=wp.data_ptr()),
           _lib.make_op(_lib.OP_CONV, dtype=_lib.BF16, mode=2, src0=xs.data_ptr(), c0=64, c1=0, ld0=16, ld1=0, up0=0, B=B, Hi=Hs, Wi=Ws, Ho=Hs, Wo=Ws,
                        kh=4, kw=1, stride=1, pad=2, w=wp.data_ptr(), Cout=64, out=out.data_ptr(), ldo=64, stats=st.data_ptr())]
    ol = _lib.OpList(ops)
    ol.run(stream)
    torch.cuda.synchronize()
    ref = F.conv2d(x.bfloat16().float(), w.bfloat16().float(), stride=2, padding=3).permute(0, 2, 3, 1)
    err = ((out.float() - ref).norm() / ref.norm()).item()
    serr = ((st[0] - ref.double().sum((0, 1, 2))).norm() / ref.double().sum((0, 1, 2)).norm()).item()
    ms = None
    if reps:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ol2 = _lib.OpList([ops[2]] * reps)
        ol2.run(stream); torch.cuda.synchronize()
        e0.record(); ol2.run(stream); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
    print(f"B={B} {H}x{W}: rel err {err:.3e} stats err {serr:.3e}" + (f"  conv {ms * 1e3:.1f} us" if ms else ""), flush=True)

for c in [(2, 64, 64), (3, 32, 32), (1, 128, 128), (2, 256, 256), (5, 64, 64)]:
    try:
        run(*c)
    except Exception as ex:
        print("case", c, "FAILED:", str(ex)[:300], flush=True)
for c in [(256, 64, 64), (64, 128, 128)]:
    try:
        run(*c, reps=20)
    except Exception as ex:
        print("case", c, "FAILED:", str(ex)[:300], flush=True)
print("device error flag:", _lib.load().d3fk_device_error_flag())
