"""Run one representative bf16 conv (layer2-like, 128->128 ch 3x3 on 16x16, B=256 => M=65536, N=128, K=1152) a few times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from denoising_diffusion_deep_fake_b200 import _lib
_lib.init(0)
B, C, H, Cout = 256, 128, 16, 128
if len(sys.argv) > 4:
    B, C, H, Cout = (int(a) for a in sys.argv[1:5])
dev = "cuda:0"; bf = torch.bfloat16
src0 = torch.randn(B, H, H, C, device=dev).to(bf)
w = (torch.randn(Cout, 9 * C, device=dev) * 0.03).to(bf)
out = torch.zeros(B, H, H, Cout, device=dev, dtype=bf)
stats = torch.zeros(2, Cout, device=dev, dtype=torch.float64)
op = _lib.make_op(_lib.OP_CONV, dtype=_lib.BF16, mode=0, src0=src0.data_ptr(), c0=C, ld0=C, B=B, Hi=H, Wi=H, Ho=H, Wo=H, kh=3, kw=3,
                  stride=1, pad=1, w=w.data_ptr(), Cout=Cout, out=out.data_ptr(), ldo=Cout, stats=stats.data_ptr())
s = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    _lib.run_single(op, s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    _lib.run_single(op, s)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
fl = 2.0 * B * H * H * Cout * 9 * C
print(f"conv M={B*H*H} N={Cout} K={9*C}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TF/s")
