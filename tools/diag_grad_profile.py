"""Where does the gradient error of one training step enter?  Per-tensor relative error (against the float64 oracle) in
BACKWARD order for d3fk fp32 / bf16 and for torch's own fp32 / bf16-autocast runs of the oracle, plus the conditioning of
every BatchNorm layer (min over channels of batch variance, and of variance / mean^2) on the same inputs.
    python tools/diag_grad_profile.py [B] [face seed] [noise seed]"""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oracle
import denoising_diffusion_deep_fake_b200 as d3
from tools.diag_bf16_evidence import faces, rel, torch_run, d3fk_run

DEV = "cuda:0"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
fs = int(sys.argv[2]) if len(sys.argv) > 2 else 21
ns = int(sys.argv[3]) if len(sys.argv) > 3 else 23
torch.manual_seed(0)
ref = oracle.Unet()
sd0 = {k: v.clone() for k, v in ref.state_dict().items()}
sd = {k: v.cpu() for k, v in oracle.short_training_run(ref, sd0, steps=150, device=DEV).items()}
g = torch.Generator(device=DEV).manual_seed(fs)
x0 = (torch.nn.functional.avg_pool2d(0.5 * torch.randn(B, 3, 64, 64, generator=g, device=DEV), 5, 1, 2) * 2.5).clamp(-1, 1).contiguous()
g = torch.Generator(device=DEV).manual_seed(ns)
noise = torch.randn(x0.shape, generator=g, device=DEV)
y = torch.rand((B, 1, 1, 1), generator=g, device=DEV)
noisy = d3.q_sample(x0, 5.0, noise=noise, y=y)
# BN conditioning on these inputs (fp64 oracle, train mode)
m64 = copy.deepcopy(ref); m64.load_state_dict(sd); m64 = m64.double().to(DEV).train()
cond = {}
def hook(name):
    def f(mod, inp, out):
        x = inp[0]
        var = x.var(dim=(0, 2, 3), unbiased=False); mean = x.mean(dim=(0, 2, 3))
        cond[name] = (var.min().item(), (var / (mean * mean + 1e-300)).min().item(), (mod.weight.abs() / torch.sqrt(var + 1e-5)).max().item())
    return f
hs = [mod.register_forward_hook(hook(n)) for n, mod in m64.named_modules() if isinstance(mod, torch.nn.BatchNorm2d)]
with torch.no_grad():
    m64(noisy.double())
for h in hs: h.remove()
print("BN layers by smallest batch variance (name, min var, min var/mean^2, max gamma*invstd):")
for n, v in sorted(cond.items(), key=lambda kv: kv[1][0])[:10]:
    print(f"  {n:36s} {v[0]:.3e} {v[1]:.3e} {v[2]:.3e}")
p64, g64, names = torch_run(ref, sd, torch.float64, False, noisy, x0)
runs = {}
runs["torch fp32"] = torch_run(ref, sd, torch.float32, False, noisy, x0)[1]
runs["d3fk fp32"] = d3fk_run(sd, "fp32", noisy, x0)[1]
runs["torch bf16ac"] = torch_run(ref, sd, torch.float32, True, noisy, x0)[1]
runs["d3fk bf16"] = d3fk_run(sd, "bf16", noisy, x0)[1]
from denoising_diffusion_deep_fake_b200.plan import backward_param_order
order = backward_param_order()
print(f"\nper-tensor relative error in backward order (B={B}, face seed {fs}, noise seed {ns})")
print(f"{'tensor':44s} " + " ".join(f"{k:>13s}" for k in runs))
for n in order:
    if n.endswith(".weight") and ("conv" in n or "downsample.0" in n or "segmentation" in n) and g64[n].dim() == 4:
        print(f"{n:44s} " + " ".join(f"{rel(runs[k][n], g64[n]):13.3e}" for k in runs))
for k in runs:
    flat = torch.cat([runs[k][n].flatten().double() for n in order]); flat64 = torch.cat([g64[n].flatten() for n in order])
    print(f"{k}: arena rel {rel(flat, flat64):.3e}")
