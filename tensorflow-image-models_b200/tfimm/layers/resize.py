"""Cold-path (load-time) resampling of position embeddings.

``interpolate_pos_embeddings`` mirrors tfimm/layers/transformers.py:13-47, which calls
``tf.image.resize(method="bicubic")``.  TF2's bicubic resize is the Keys cubic kernel with
A = -0.5, half-pixel centres, no antialiasing, and -- unlike OpenCV / PyTorch -- taps that fall
outside the image are dropped and the remaining weights renormalised
(tensorflow/core/kernels/image/scale_and_translate_op.cc, ComputeSpans).  It runs once per
weight load on a (1, N, D) tensor, so it is plain torch on whatever device the weight lives on.
"""
from typing import Tuple

import torch


def _keys_cubic(x: torch.Tensor) -> torch.Tensor:
    a = -0.5
    x = x.abs()
    near = ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0
    far = ((a * x - 5.0 * a) * x + 8.0 * a) * x - 4.0 * a
    return torch.where(x <= 1.0, near, torch.where(x < 2.0, far, torch.zeros_like(x)))


def _resize_matrix(n_in: int, n_out: int, dtype, device) -> torch.Tensor:
    """(n_out, n_in) interpolation matrix of TF2's bicubic resize along one axis."""
    scale = n_in / n_out
    centers = (torch.arange(n_out, dtype=torch.float64, device=device) + 0.5) * scale
    src = torch.arange(n_in, dtype=torch.float64, device=device) + 0.5
    w = _keys_cubic(src[None, :] - centers[:, None])  # kernel scale 1 (no antialias)
    w = w / w.sum(dim=1, keepdim=True)
    return w.to(dtype)


def tf_bicubic_resize(images: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
    """``tf.image.resize(images, size, method="bicubic")`` for NHWC tensors."""
    _, h, w, _ = images.shape
    mh = _resize_matrix(h, size[0], images.dtype, images.device)
    mw = _resize_matrix(w, size[1], images.dtype, images.device)
    out = torch.einsum("oh,bhwc->bowc", mh, images)
    return torch.einsum("pw,bowc->bopc", mw, out)


def interpolate_pos_embeddings(pos_embed: torch.Tensor, src_grid_size, tgt_grid_size, nb_tokens: int = 0):
    """(1, nb_tokens + h*w, D) -> (1, nb_tokens + h'*w', D); token embeddings are kept as they are."""
    src_grid_size, tgt_grid_size = tuple(src_grid_size), tuple(tgt_grid_size)
    if src_grid_size == tgt_grid_size:
        return pos_embed
    grid = pos_embed[:, nb_tokens:].reshape(1, *src_grid_size, -1)
    grid = tf_bicubic_resize(grid, tgt_grid_size)
    grid = grid.reshape(1, tgt_grid_size[0] * tgt_grid_size[1], -1)
    return torch.cat((pos_embed[:, :nb_tokens], grid), dim=1)
