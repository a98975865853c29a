"""Device-side evaluation transform of the video recipe:
`Compose([Normalize(0, 255), CenterCrop((88, 88)), Normalize(0.421, 0.165)])`
(N20EMv2/video_only/train_video_ssl.py:445-457, utils.py:45-84) followed by the recipe's permute to (B, 1, T, H, W)
(`train_video_ssl.py:34`), as one kernel behind `svt_video_transform_u8`.  No CPU fallback."""
from __future__ import annotations

import torch

from ._lib import check, current_stream_ptr, lib, ptr

IMAGE_CROP_SIZE = 88
IMAGE_MEAN = 0.421
IMAGE_STD = 0.165


@torch.no_grad()
def eval_transform(frames: torch.Tensor, crop: int = IMAGE_CROP_SIZE, mean: float = IMAGE_MEAN,
                   std: float = IMAGE_STD) -> torch.Tensor:
    """frames: uint8 CUDA tensor (T, H, W) of one clip or (B, T, H, W) -> fp32 (B, 1, T, crop, crop), the input of
    `FairseqAVHubertPretrain.forward({"video": ..., "audio": None})`."""
    if frames.dtype != torch.uint8:
        raise TypeError(f"expected uint8 grey frames, got {frames.dtype}")
    if not frames.is_cuda:
        raise RuntimeError("svt_speechbrain_b200.video_transforms runs on CUDA only; no CPU fallback")
    if frames.dim() == 3:
        frames = frames.unsqueeze(0)
    if frames.dim() != 4:
        raise ValueError(f"expected (T, H, W) or (B, T, H, W), got {tuple(frames.shape)}")
    frames = frames.contiguous()
    B, T, H, W = frames.shape
    out = torch.empty(B, 1, T, crop, crop, dtype=torch.float32, device=frames.device)
    with torch.cuda.device(frames.device):
        check(lib().svt_video_transform_u8(ptr(frames), B * T, H, W, crop, mean, std, ptr(out), current_stream_ptr()))
    return out
