"""Batched inference render (reference: render_strings / binary_array_to_image,
helpers.py:20-74). The reference renders one string per forward call and converts on the host;
here all strings go through one eval forward whose GEMM epilogue already emits the uint8 pixels
(helpers.py:33 truncation), and the host only writes the 8-bpp BMP files.
"""
from __future__ import annotations

import os
from typing import Sequence

import numpy as np
import torch

from .data import encode

_BMP_HEADER_BYTES = 14 + 40 + 256 * 4   # file header + BITMAPINFOHEADER + grey palette = 1078


def grey_bmp_bytes(img: np.ndarray) -> bytes:
    """8-bit palettised BMP exactly as PIL writes a mode-'L' image (helpers.py:36,42): 1078-byte
    header, identity grey palette, bottom-up rows padded to 4 bytes, 96 dpi."""
    assert img.dtype == np.uint8 and img.ndim == 2
    h, w = img.shape
    stride = (w + 3) & ~3
    image_bytes = stride * h
    ppm = int(96 * 39.3701 + 0.5)
    header = b"BM" + (_BMP_HEADER_BYTES + image_bytes).to_bytes(4, "little") + (0).to_bytes(4, "little") \
        + _BMP_HEADER_BYTES.to_bytes(4, "little")
    info = (40).to_bytes(4, "little") + w.to_bytes(4, "little") + h.to_bytes(4, "little") \
        + (1).to_bytes(2, "little") + (8).to_bytes(2, "little") + (0).to_bytes(4, "little") \
        + image_bytes.to_bytes(4, "little") + ppm.to_bytes(4, "little") + ppm.to_bytes(4, "little") \
        + (256).to_bytes(4, "little") + (256).to_bytes(4, "little")
    palette = np.repeat(np.arange(256, dtype=np.uint8)[:, None], 4, axis=1)
    palette[:, 3] = 0
    rows = np.zeros((h, stride), dtype=np.uint8)
    rows[:, :w] = img[::-1]
    return header + info + palette.tobytes() + rows.tobytes()


def strings_to_tokens(strings: Sequence[str], max_length: int) -> torch.Tensor:
    """helpers.py:52-59: truncate to max_length, ord(), pad with token 0."""
    return encode([s[:max_length] for s in strings], max_length)


@torch.no_grad()
def render_batch_u8(model, strings: Sequence[str], device, batch_size: int = 4096) -> torch.Tensor:
    """uint8 [N,H,W] on the device, rendered in batches (no file I/O)."""
    tokens = strings_to_tokens(strings, model.max_length).to(device)
    outs = []
    for i in range(0, tokens.shape[0], batch_size):
        outs.append(model.render_u8(tokens[i:i + batch_size]))
    return torch.cat(outs, dim=0) if len(outs) > 1 else outs[0]


def render_strings(model, strings, output_dir, sheet_height, sheet_width, device):
    """Render a list of strings as BMP images -- same signature, file names, truncation warning and
    summary line as helpers.py:46-74. Like the reference it does not switch the model to eval():
    callers do (model.py:314, helpers.py:103)."""
    os.makedirs(output_dir, exist_ok=True)
    strings = list(strings)
    for i, s in enumerate(strings):
        if len(s) > model.max_length:
            strings[i] = s[:model.max_length]
            print(f"Warning: String truncated to {model.max_length} characters: {strings[i]}")
    if strings:
        was_training = model.training
        if was_training:
            # helpers.py:64 would run a dropout forward here; that only happens if a caller
            # forgot model.eval(). Keep the quirk observable but do not add RNG files to disk:
            with torch.no_grad():
                sheets = (model(strings_to_tokens(strings, model.max_length).to(device)) * 255).to(torch.uint8)
        else:
            sheets = render_batch_u8(model, strings, device)
        host = sheets.cpu().numpy()
        for idx in range(len(strings)):
            with open(f"{output_dir}/string_{idx}.bmp", "wb") as f:
                f.write(grey_bmp_bytes(host[idx]))
    print(f"Saved {len(strings)} rendered strings to {output_dir}/")
