"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement of the reference's video-frame pre / post-processing around `predict_fake`
(d3f/train_deep_fake/lit_module.py:272-300), batched over a leading frame axis.  cv2.cvtColor(BGR2RGB / RGB2BGR)
is a reversal of the channel axis; everything else is the reference's own torch arithmetic, op for op."""
import numpy as np
import torch


def cv2_to_tensor_normalised(frames_bgr, mean, std):
    """lit_module.py:272-283.  frames_bgr: uint8 [N,H,W,3] (BGR); mean / std: 3 floats (RGB order).  -> fp32 [N,3,H,W]."""
    mean = torch.tensor(mean)                                     # :261-262 (float32)
    std = torch.tensor(std)
    image_rgb = np.ascontiguousarray(np.asarray(frames_bgr)[..., ::-1])   # cv2.COLOR_BGR2RGB (:274)
    tensor = torch.from_numpy(image_rgb).float()                  # :276
    tensor = tensor.permute(0, 3, 1, 2).contiguous()              # hwc to chw (:278)
    tensor -= mean.reshape(3, 1, 1) * 255                         # :280
    tensor /= std.reshape(3, 1, 1) * 255                          # :281
    return tensor


def tensor_cv2_to_denormalised(tensor, mean, std):
    """lit_module.py:285-300.  tensor: fp32 [N,3,H,W] (RGB) -> uint8 [N,H,W,3] (BGR)."""
    mean = torch.tensor(mean)
    std = torch.tensor(std)
    tensor = tensor.clone()
    tensor *= std.reshape(3, 1, 1) * 255                          # :288
    tensor += mean.reshape(3, 1, 1) * 255                         # :289
    tensor = tensor.permute(0, 2, 3, 1)                           # chw to hwc (:291)
    tensor = tensor.int()                                         # :293 (truncates toward zero)
    tensor = tensor.clamp(0, 255)                                 # :294
    image_rgb = tensor.cpu().numpy().astype(np.uint8)             # :296
    return np.ascontiguousarray(image_rgb[..., ::-1])             # cv2.COLOR_RGB2BGR (:298)
