"""Where a training step's device time goes, stream by stream: CUDA events on the main stream after the forward, the
loss, each backward segment, the weight-gradient join and the update stream (StepOverlap), averaged over steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from denoising_diffusion_deep_fake_b200 import _lib
from denoising_diffusion_deep_fake_b200.train import DenoiserModule, FlatAdam
from denoising_diffusion_deep_fake_b200 import plan as planmod

dev = torch.device("cuda:0")
mod = DenoiserModule(encoder_name="resnet34", learning_rate=0.02, noise_exponential_sampling_lambda=5,
                     cosine_scheduler_max_epoch=100, precision="bf16").to(dev).train()
mod.configure_optimizers(fused=True)
x = torch.randn(256, 3, 64, 64, device=dev).clamp(-1, 1)
for _ in range(5):
    mod.training_step(x)
torch.cuda.synchronize()

marks = []          # (label, event) of the current step
def mark(label, stream=None):
    ev = torch.cuda.Event(enable_timing=True)
    ev.record(stream or torch.cuda.current_stream())
    marks.append((label, ev))

orig_run_backward = planmod.UnetPlan.run_backward
def run_backward(self, dy, stream, after_segment=None):
    mark("loss+loss-bwd")
    def hook(i):
        mark(f"bwd seg {i} (main chain)")
        if after_segment is not None:
            after_segment(i)
    from denoising_diffusion_deep_fake_b200._lib import op_params
    op_params(self.bwd_segments[0].array[self.dy_op_index]).src = dy.data_ptr()
    for i, seg in enumerate(self.bwd_segments):
        seg.run(stream, join=False)
        hook(i)
    _lib.side_stream_join(stream)
    mark("wgrad join (main waits for side streams)")
planmod.UnetPlan.run_backward = run_backward

orig_fwd = planmod.UnetPlan.run_forward
def run_forward(self, x_, y, stream):
    mark("q_sample+pack")
    orig_fwd(self, x_, y, stream)
    mark("forward")
planmod.UnetPlan.run_forward = run_forward

N = 20
acc = {}
order = []
for it in range(N):
    marks.clear()
    mark("start")
    mod.training_step(x)
    mark("update stream joined + tail (Adam/pack unless overlapped)")
    torch.cuda.synchronize()
    for (l0, e0), (l1, e1) in zip(marks[:-1], marks[1:]):
        acc[l1] = acc.get(l1, 0.0) + e0.elapsed_time(e1)
        if l1 not in order:
            order.append(l1)
    acc["TOTAL"] = acc.get("TOTAL", 0.0) + marks[0][1].elapsed_time(marks[-1][1])
print("overlap:", mod.allreduce is not None)
for l in order + ["TOTAL"]:
    print(f"{acc[l] / N:8.3f} ms  {l}")
