"""Print the launch geometry of every conv / wgrad of one training step (D3FK_VERBOSE=1)."""
import os, sys
os.environ["D3FK_VERBOSE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from denoising_diffusion_deep_fake_b200.train import DenoiserModule
dev = torch.device("cuda:0")
mod = DenoiserModule(encoder_name="resnet34", learning_rate=0.02, noise_exponential_sampling_lambda=5,
                     cosine_scheduler_max_epoch=100, precision="bf16").to(dev).train()
mod.configure_optimizers(fused=True)
x = torch.randn(256, 3, 64, 64, device=dev).clamp(-1, 1)
mod.training_step(x)
torch.cuda.synchronize()
