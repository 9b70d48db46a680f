"""Data conventions of the hot path's callers (helpers.py:107-181, generate_font.ts:164-199) and
the synthetic workloads bench.py / the tests use when the real FiraCode bitmaps are not on disk
(bun + node-canvas cannot run offline).

Everything here is host-side setup; the device-resident layout it produces is what the kernels
consume: tokens int64 [N, Lmax] (0-padded) and targets uint8 [N, H, W] (255 = white).
"""
from __future__ import annotations

import math
import os
from typing import List, Sequence, Tuple

import numpy as np
import torch

LCG_MUL, LCG_ADD, LCG_MOD = 1664525, 1013904223, 4294967296


def seeded_text(seed: int, min_len: int = 10, max_len: int = 100) -> str:
    """The string generate_font.ts produces for one sample (its LCG, word and space rules,
    generate_font.ts:164-199). JS computes floor(seed / 2**32 * n) in doubles; seed < 2**32 and
    n <= 91 make that exactly (seed * n) >> 32."""
    state = seed

    def draw(n: int) -> int:
        nonlocal state
        state = (state * LCG_MUL + LCG_ADD) % LCG_MOD
        return (state * n) >> 32

    remaining = draw(max_len - min_len + 1) + min_len
    parts: List[str] = []
    while remaining > 0:
        wl = min(draw(10) + 1, remaining)
        parts.append("".join(chr(65 + draw(26)) for _ in range(wl)))
        remaining -= wl
        if remaining > 0:
            parts.append(" ")
            remaining -= 1
    return "".join(parts)


def dataset_texts(n: int, first_seed: int = 42) -> List[str]:
    """Line i of train_input/data.txt (generate_font.ts:203-216): seed = i + 42."""
    return [seeded_text(first_seed + i) for i in range(n)]


def encode(strings: Sequence[str], width: int) -> torch.Tensor:
    """ord() per character, right-padded with 0 to `width` (helpers.py:57-59,163-177)."""
    arr = np.zeros((len(strings), width), dtype=np.int64)
    for i, s in enumerate(strings):
        codes = [ord(c) for c in s[:width]]          # any code point, like the reference
        arr[i, : len(codes)] = codes
    return torch.from_numpy(arr)


def encode_with_font(strings: Sequence[str], font_ids: Sequence[int], width: int,
                     n_chars: int = 128) -> torch.Tensor:
    """Multi-font conditioning (BASELINE config 3; an extension -- the reference has one font and no
    conditioning input): the font is a CONTROL TOKEN. Position 0 holds token n_chars + font_id, the
    characters follow from position 1, zero padding as before; the model is built with
    vocab = n_chars + n_fonts and max_length = longest string + 1. The token's embedding row plays
    the role of a font embedding and reaches every position through the attention layer, and
    because it is just a vocabulary row the forward / backward / optimizer kernels, the oracle and
    the data-parallel path apply unchanged (SURVEY 8d proposed adding a font vector to every token
    embedding instead; that needs a thirteenth parameter tensor in the checkpoint layout)."""
    if len(strings) != len(font_ids):
        raise ValueError("one font id per string")
    body = encode(strings, width - 1)
    tokens = torch.zeros((len(strings), width), dtype=torch.int64)
    tokens[:, 1:] = body
    tokens[:, 0] = torch.as_tensor(list(font_ids), dtype=torch.int64) + n_chars
    return tokens


def synthetic_sheets(strings: Sequence[str], height: int = 80, width: int = 240,
                     seed: int = 1234) -> np.ndarray:
    """uint8 [N,H,W] stand-ins for the FiraCode renders: white (255) with ~5 % ink at four grey
    levels in the rows the wrapped text would occupy (33 chars/line, 14.4 px line pitch)."""
    rng = np.random.default_rng(seed)
    n = len(strings)
    out = np.full((n, height, width), 255, dtype=np.uint8)
    levels = np.array([0, 64, 128, 192], dtype=np.uint8)
    for i, s in enumerate(strings):
        rows = min(height, max(1, math.ceil(len(s) / 33) * 14))
        ink = rng.random((rows, width)) < 0.05
        vals = levels[rng.integers(0, 4, size=(rows, width))]
        out[i, :rows][ink] = vals[ink]
    return out


def synthetic_batch(n: int, max_length: int = 100, height: int = 80, width: int = 240,
                    seed: int = 1234, first_seed: int = 42) -> Tuple[torch.Tensor, torch.Tensor]:
    """(tokens int64 [n, max_length], targets uint8 [n, H, W]) -- the bench workload."""
    texts = dataset_texts(n, first_seed)
    return encode(texts, max_length), torch.from_numpy(synthetic_sheets(texts, height, width, seed))


def fast_synthetic_batch(n: int, max_length: int = 100, height: int = 80, width: int = 240,
                         seed: int = 1234) -> Tuple[torch.Tensor, torch.Tensor]:
    """Vectorised variant for large n (same distributions, different draws): lengths U{10..100},
    letters A-Z with ~15 % spaces, 5 % ink in the text rows."""
    g = torch.Generator().manual_seed(seed)
    lengths = torch.randint(10, 101, (n,), generator=g).clamp(max=max_length)
    letters = torch.randint(65, 91, (n, max_length), generator=g)
    spaces = torch.rand((n, max_length), generator=g) < 0.148
    tokens = torch.where(spaces, torch.full_like(letters, 32), letters)
    pos = torch.arange(max_length).unsqueeze(0)
    tokens = torch.where(pos < lengths.unsqueeze(1), tokens, torch.zeros_like(tokens))
    rows = ((lengths + 32) // 33 * 14).clamp(max=height)
    ink = torch.rand((n, height, width), generator=g) < 0.05
    ink &= (torch.arange(height).view(1, height, 1) < rows.view(n, 1, 1))
    vals = (torch.randint(0, 4, (n, height, width), generator=g) * 64).to(torch.uint8)
    targets = torch.where(ink, vals, torch.full_like(vals, 255))
    return tokens.long(), targets


def targets_as_u8(targets: torch.Tensor):
    """The TensorDataset of helpers.py:177-181 stores fp32 k/255 values (they came from 8-bit
    bitmaps, helpers.py:118-121). Returns the lossless uint8 form, or None if some value is not
    exactly k/255 (then the fp32 target path of the kernels is used)."""
    if targets.dtype == torch.uint8:
        return targets
    q = torch.round(targets * 255.0)
    back = q / 255.0
    if torch.equal(back, targets) and float(q.min()) >= 0 and float(q.max()) <= 255:
        return q.to(torch.uint8)
    return None


def read_bmp_grey(path: str) -> np.ndarray:
    """24-bit (generate_font.ts:6-62: BGR, top-down when height < 0, rows padded to 4 bytes) or
    8-bit palettised BMP -> uint8 grey [H,W] with PIL's convert('L') weights
    (L = (R*19595 + G*38470 + B*7471 + 0x8000) >> 16)."""
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:2] != b"BM":
        raise ValueError(f"{path}: not a BMP file")
    off = int.from_bytes(raw[10:14], "little")
    w = int.from_bytes(raw[18:22], "little", signed=True)
    h = int.from_bytes(raw[22:26], "little", signed=True)
    bpp = int.from_bytes(raw[28:30], "little")
    top_down = h < 0
    h = abs(h)
    if bpp == 24:
        stride = (w * 3 + 3) & ~3
        px = np.frombuffer(raw, dtype=np.uint8, count=stride * h, offset=off).reshape(h, stride)
        bgr = px[:, : w * 3].reshape(h, w, 3).astype(np.uint32)
        grey = (bgr[..., 2] * 19595 + bgr[..., 1] * 38470 + bgr[..., 0] * 7471 + 0x8000) >> 16
        grey = grey.astype(np.uint8)
    elif bpp == 8:
        stride = (w + 3) & ~3
        pal = np.frombuffer(raw, dtype=np.uint8, count=256 * 4, offset=54).reshape(256, 4).astype(np.uint32)
        lut = ((pal[:, 2] * 19595 + pal[:, 1] * 38470 + pal[:, 0] * 7471 + 0x8000) >> 16).astype(np.uint8)
        idx = np.frombuffer(raw, dtype=np.uint8, count=stride * h, offset=off).reshape(h, stride)[:, :w]
        grey = lut[idx]
    else:
        raise ValueError(f"{path}: unsupported BMP depth {bpp}")
    return grey if top_down else grey[::-1].copy()


def load_string_dataset_u8(data_dir: str = "train_input", num_samples: int = 50000,
                           sheet_height: int = 80, sheet_width: int = 240):
    """Same inputs and errors as helpers.py:125-181, but keeps the lossless uint8 sheets
    (2.9 GB for 150k samples instead of 11.5 GB fp32). Returns (tokens int64 [N,Lmax], uint8 [N,H,W])."""
    strings_path = os.path.join(data_dir, "data.txt")
    with open(strings_path, "r") as f:
        strings = f.read().splitlines()
    if len(strings) < num_samples:
        raise ValueError(f"Not enough strings in {strings_path}. Expected {num_samples}, got {len(strings)}")
    targets = np.zeros((num_samples, sheet_height, sheet_width), dtype=np.uint8)
    for i in range(num_samples):
        image_path = os.path.join(data_dir, f"{i + 1}.bmp")
        if not os.path.exists(image_path):
            raise FileNotFoundError(f"Image file not found: {image_path}")
        targets[i] = read_bmp_grey(image_path)
    strings = strings[:num_samples]
    max_len = max(len(s) for s in strings)
    return encode(strings, max_len), torch.from_numpy(targets)


def load_multifont_dataset_u8(data_dir: str = "train_input", num_samples: int = 50000,
                              sheet_height: int = 80, sheet_width: int = 240, n_chars: int = 128):
    """A fontgen multi-font set (data.txt, <i>.bmp, fonts.txt with one font index per line) ->
    (tokens int64 [N, Lmax + 1] with the font control token in column 0, uint8 sheets [N,H,W],
    number of fonts)."""
    tokens, targets = load_string_dataset_u8(data_dir, num_samples, sheet_height, sheet_width)
    with open(os.path.join(data_dir, "fonts.txt")) as f:
        fonts = [int(x) for x in f.read().split()][:num_samples]
    if len(fonts) < num_samples:
        raise ValueError(f"Not enough font ids in {data_dir}/fonts.txt. Expected {num_samples}, got {len(fonts)}")
    out = torch.zeros((num_samples, tokens.shape[1] + 1), dtype=torch.int64)
    out[:, 1:] = tokens
    out[:, 0] = torch.tensor(fonts, dtype=torch.int64) + n_chars
    return out, targets, max(fonts) + 1


class HostBatchFeeder:
    """Streams (tokens, uint8 sheets) batches from pinned host memory to the device on a copy
    stream, one batch ahead of the training step that consumes them, so the 20 MB per-batch H2D
    copy (model.py:295-296 does it synchronously in front of every step) hides under the previous
    step's kernels. Double buffered: `get(i)` hands out the device tensors of batch i and starts
    the copy of batch i+1; the compute stream is made to wait for the copy it consumes, and the
    copy stream for the step that last read the buffer it overwrites."""

    def __init__(self, tokens_host: torch.Tensor, targets_host: torch.Tensor, batch: int,
                 device: torch.device):
        if not (tokens_host.is_pinned() and targets_host.is_pinned()):
            raise ValueError("HostBatchFeeder needs pinned host tensors (tensor.pin_memory())")
        self.tok_h, self.tgt_h, self.batch, self.device = tokens_host, targets_host, batch, device
        self.n_batches = tokens_host.shape[0] // batch
        self.copy_stream = torch.cuda.Stream(device=device)
        self.bufs = [(torch.empty((batch,) + tuple(tokens_host.shape[1:]), dtype=tokens_host.dtype, device=device),
                      torch.empty((batch,) + tuple(targets_host.shape[1:]), dtype=targets_host.dtype, device=device))
                     for _ in range(2)]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]     # copy into buffer k finished
        self.consumed = [None, None]                               # last step reading buffer k finished
        self.in_flight = None                                      # batch index being / already copied ahead
        self.h2d_bytes_per_batch = (batch * tokens_host[0].numel() * tokens_host.element_size()
                                    + batch * targets_host[0].numel() * targets_host.element_size())

    def _start_copy(self, i: int):
        k = i % 2
        lo = (i % self.n_batches) * self.batch
        with torch.cuda.stream(self.copy_stream):
            if self.consumed[k] is not None:
                self.copy_stream.wait_event(self.consumed[k])
            self.bufs[k][0].copy_(self.tok_h[lo:lo + self.batch], non_blocking=True)
            self.bufs[k][1].copy_(self.tgt_h[lo:lo + self.batch], non_blocking=True)
            self.ready[k].record(self.copy_stream)
        self.in_flight = i

    def get(self, i: int):
        """Device (tokens, targets) of batch i for the current stream; prefetches batch i+1."""
        if self.in_flight != i:
            self._start_copy(i)
        k = i % 2
        torch.cuda.current_stream(self.device).wait_event(self.ready[k])
        self._start_copy(i + 1)
        return self.bufs[k]

    def done(self, i: int):
        """Call after the last kernel reading batch i has been enqueued on the current stream."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.consumed[i % 2] = ev
