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


def strings_to_tokens(strings: Sequence[str], max_length: int, font_ids=None) -> torch.Tensor:
    """helpers.py:52-59: truncate to max_length, ord(), pad with token 0. With font_ids (multi-font
    models, data.encode_with_font) position 0 carries the font control token."""
    if font_ids is not None:
        from .data import encode_with_font
        return encode_with_font([s[:max_length - 1] for s in strings], font_ids, max_length)
    return encode([s[:max_length] for s in strings], max_length)


def check_token_range(model, tokens: torch.Tensor):
    """nn.Embedding raises IndexError for an id outside the table (model.py:167, e.g. a character
    >= 128 with the reference's 128-row vocabulary); so does the render path, on the host, before
    anything is launched."""
    vocab = model.embedding.num_embeddings
    if tokens.numel() and (int(tokens.max()) >= vocab or int(tokens.min()) < 0):
        raise IndexError("index out of range in self")


@torch.no_grad()
def render_batch_u8(model, strings: Sequence[str], device, batch_size: int = 4096) -> torch.Tensor:
    """uint8 [N,H,W] on the device, rendered in batches (no file I/O)."""
    tokens = strings_to_tokens(strings, model.max_length)
    check_token_range(model, tokens)
    tokens = tokens.to(device)
    outs = []
    for i in range(0, tokens.shape[0], batch_size):
        outs.append(model.render_u8(tokens[i:i + batch_size]))
    return torch.cat(outs, dim=0) if len(outs) > 1 else outs[0]


class RenderPipeline:
    """Batched render into pinned HOST memory: batch i's uint8 sheets travel device -> host on a
    copy stream while batch i+1 is rendered (two device buffers, CUDA events both ways), so the
    19.2 KB/glyph of output crosses PCIe under the GEMM instead of after it. The reference
    (helpers.py:62-68) does one forward, one `.cpu()` and one PIL call per string."""

    def __init__(self, model, device, batch_size: int = 4096):
        self.model, self.device, self.batch = model, torch.device(device), batch_size
        shape = (batch_size, model.sheet_height, model.sheet_width)
        self.dev = [torch.empty(shape, dtype=torch.uint8, device=self.device) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.rendered = [torch.cuda.Event() for _ in range(2)]
        self.copied = [None, None]

    @torch.no_grad()
    def render_to_host(self, tokens: torch.Tensor, out: torch.Tensor = None, on_batch=None) -> torch.Tensor:
        """tokens: int64 [N, L] (host, ideally pinned, or device). Returns uint8 [N,H,W] in pinned
        host memory (`out` if given). on_batch(lo, hi, done_event), if given, is called after the
        D2H copy of sheets [lo, hi) has been enqueued; done_event fires when they are on the host."""
        n = tokens.shape[0]
        H, W = self.model.sheet_height, self.model.sheet_width
        if out is None:
            out = torch.empty((n, H, W), dtype=torch.uint8).pin_memory()
        main = torch.cuda.current_stream(self.device)
        for i, lo in enumerate(range(0, n, self.batch)):
            hi, k = min(n, lo + self.batch), i % 2
            x = tokens[lo:hi].to(self.device, non_blocking=True)
            if self.copied[k] is not None:
                main.wait_event(self.copied[k])            # buffer k has left for the host
            y = self.dev[k][: hi - lo]
            self.model.render_u8(x, out=y)
            self.rendered[k].record(main)
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(self.rendered[k])
                out[lo:hi].copy_(y, non_blocking=True)
                done = torch.cuda.Event()
                done.record(self.copy_stream)
            self.copied[k] = done
            if on_batch is not None:
                on_batch(lo, hi, done)
        main.wait_stream(self.copy_stream)
        return out


def bmp_file_block(sheets: np.ndarray) -> np.ndarray:
    """[n, 1078 + H*stride] uint8: the complete file image of every sheet (same bytes as
    grey_bmp_bytes), built with three array copies instead of n Python loops."""
    assert sheets.dtype == np.uint8 and sheets.ndim == 3
    n, h, w = sheets.shape
    stride = (w + 3) & ~3
    head = np.frombuffer(grey_bmp_bytes(np.zeros((h, w), dtype=np.uint8))[:_BMP_HEADER_BYTES], dtype=np.uint8)
    block = np.zeros((n, _BMP_HEADER_BYTES + h * stride), dtype=np.uint8)
    block[:, :_BMP_HEADER_BYTES] = head
    block[:, _BMP_HEADER_BYTES:].reshape(n, h, stride)[:, :, :w] = sheets[:, ::-1, :]
    return block


def write_bmp_files(sheets: np.ndarray, output_dir: str, first_index: int = 0, threads: int = 8,
                    name: str = "string_{}.bmp"):
    """Writes sheets[i] as output_dir/string_{first_index+i}.bmp (helpers.py:66-68 naming) from a
    small thread pool (file writes release the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    block = bmp_file_block(sheets)

    def one(i):
        with open(os.path.join(output_dir, name.format(first_index + i)), "wb") as f:
            f.write(memoryview(block[i]))

    if len(block) <= 8 or threads <= 1:
        for i in range(len(block)):
            one(i)
    else:
        with ThreadPoolExecutor(max_workers=threads) as pool:
            list(pool.map(one, range(len(block))))


def render_strings(model, strings, output_dir, sheet_height, sheet_width, device, batch_size: int = 4096,
                   font_ids=None):
    """Render a list of strings as BMP images -- same signature, file names, truncation warning and
    summary line as helpers.py:46-74, for any number of strings: batches go through
    RenderPipeline, and each batch's files are written while the next one renders. Like the
    reference it does not switch the model to eval(): callers do (model.py:314, helpers.py:103)."""
    os.makedirs(output_dir, exist_ok=True)
    strings = list(strings)
    table_fonts = font_ids is not None and getattr(model, "n_fonts", 0) > 0   # font_embedding table model
    if table_fonts:
        # multi-font model with a font_embedding table: ordinary tokens, the font goes in beside them
        for i, s in enumerate(strings):
            if len(s) > model.max_length:
                strings[i] = s[:model.max_length]
                print(f"Warning: String truncated to {model.max_length} characters: {strings[i]}")
        if strings:
            tokens = strings_to_tokens(strings, model.max_length)
            check_token_range(model, tokens)
            fonts = torch.as_tensor(list(font_ids), dtype=torch.int32)
            with torch.no_grad():
                for lo in range(0, len(strings), batch_size):
                    hi = min(len(strings), lo + batch_size)
                    sheets = model.render_u8(tokens[lo:hi].to(device), font_ids=fonts[lo:hi])
                    write_bmp_files(sheets.cpu().numpy(), output_dir, first_index=lo)
        print(f"Saved {len(strings)} rendered strings to {output_dir}/")
        return
    limit = model.max_length - (1 if font_ids is not None else 0)
    for i, s in enumerate(strings):
        if len(s) > limit:
            strings[i] = s[:limit]
            print(f"Warning: String truncated to {limit} characters: {strings[i]}")
    if strings:
        tokens = strings_to_tokens(strings, model.max_length, font_ids)
        check_token_range(model, tokens)
        if model.training:
            # helpers.py:64 would run a dropout forward here; that only happens if a caller
            # forgot model.eval(). Keep the quirk observable:
            with torch.no_grad():
                sheets = (model(tokens.to(device)) * 255).to(torch.uint8)
            write_bmp_files(sheets.cpu().numpy(), output_dir)
        else:
            pipe = RenderPipeline(model, device, min(batch_size, len(strings)))
            pending = []
            host = pipe.render_to_host(tokens.pin_memory() if len(strings) > 64 else tokens,
                                       on_batch=lambda lo, hi, ev: pending.append((lo, hi, ev)))
            arr = host.numpy()
            for lo, hi, ev in pending:       # batch k's files are written while later copies land
                ev.synchronize()
                write_bmp_files(arr[lo:hi], output_dir, first_index=lo)
    print(f"Saved {len(strings)} rendered strings to {output_dir}/")
