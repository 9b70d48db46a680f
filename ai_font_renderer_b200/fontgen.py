"""Offline stand-in for the reference's dataset generator (generate_font.ts, run with bun +
node-canvas, neither of which exists offline): the same strings (its LCG, generate_font.ts:164-199,
restated in data.seeded_text), the same layout rules (12 px font, greedy word wrap at the sheet
width, baseline of line k at (k+1) * 14.4 px, black on white, generate_font.ts:75-97,112-130), the
same files (train_input/data.txt, 1-based <i>.bmp as 24-bit BGR top-down bitmaps with 4-byte row
padding, dataset_metadata.txt; generate_font.ts:6-62,203-239) -- but rasterised by Pillow's
FreeType instead of Cairo, so anti-aliased edge pixels differ from the reference's bitmaps
(parity at this boundary is unpinned; training targets are inputs to the hot path).

    python -m ai_font_renderer_b200.fontgen --font FiraCode-Retina.ttf --samples 1000 --out train_input
    python -m ai_font_renderer_b200.fontgen --font FiraCode-Retina.ttf --font Montserrat-Regular.ttf ...

Several --font arguments give the multi-font set of BASELINE config 3: sample i is drawn with font
i mod n, and fonts.txt holds the font index of every sample.
"""
from __future__ import annotations

import argparse
import os
from typing import List, Optional, Sequence

import numpy as np

from .data import dataset_texts

FONT_SIZE = 12
SHEET_WIDTH, SHEET_HEIGHT = 240, 80
LINE_HEIGHT = FONT_SIZE * 1.2


def load_font(path: Optional[str], size: int = FONT_SIZE):
    from PIL import ImageFont
    if path and os.path.exists(path):
        return ImageFont.truetype(path, size)
    return ImageFont.load_default()


def wrap_text(font, text: str, max_width: float) -> List[str]:
    """Greedy word wrap on single spaces; a word wider than the sheet stays on its own line
    (generate_font.ts:75-97)."""
    lines, current = [], ""
    for word in text.split(" "):
        trial = f"{current} {word}" if current else word
        if font.getlength(trial) > max_width and current:
            lines.append(current)
            current = word
        else:
            current = trial
    if current:
        lines.append(current)
    return lines


def render_sheet(font, text: str, height: int = SHEET_HEIGHT, width: int = SHEET_WIDTH) -> np.ndarray:
    """uint8 grey [H,W]: 255 = white paper, 0 = ink; line k's baseline at (k+1) * 14.4 px."""
    from PIL import Image, ImageDraw
    img = Image.new("L", (width, height), 255)
    draw = ImageDraw.Draw(img)
    for k, line in enumerate(wrap_text(font, text, width)):
        draw.text((0, (k + 1) * LINE_HEIGHT), line, fill=0, font=font, anchor="ls")
    return np.asarray(img, dtype=np.uint8)


def bmp24_topdown_bytes(grey: np.ndarray) -> bytes:
    """24-bit BGR, top-down (negative height), rows padded to 4 bytes, 54-byte header, zero
    resolution / colour counts: the file generate_font.ts:6-62 writes for an R = G = B image."""
    h, w = grey.shape
    row = (w * 3 + 3) // 4 * 4
    size = 54 + row * h
    le = lambda v, n, signed=False: int(v).to_bytes(n, "little", signed=signed)
    header = (b"BM" + le(size, 4) + le(0, 4) + le(54, 4) + le(40, 4) + le(w, 4, True) + le(-h, 4, True)
              + le(1, 2) + le(24, 2) + le(0, 4) + le(row * h, 4) + le(0, 4, True) + le(0, 4, True)
              + le(0, 4) + le(0, 4))
    px = np.zeros((h, row), dtype=np.uint8)
    px[:, : w * 3] = np.repeat(grey, 3, axis=1)
    return header + px.tobytes()


def generate_dataset(out_dir: str, num_samples: int, font_paths: Sequence[Optional[str]] = (None,),
                     height: int = SHEET_HEIGHT, width: int = SHEET_WIDTH, first_seed: int = 42,
                     quiet: bool = False) -> List[str]:
    """Writes out_dir/{data.txt, 1.bmp .. N.bmp, dataset_metadata.txt[, fonts.txt]}; returns the
    strings. helpers.load_string_dataset(out_dir, N) reads the result."""
    os.makedirs(out_dir, exist_ok=True)
    fonts = [load_font(p) for p in font_paths]
    texts = dataset_texts(num_samples, first_seed)
    with open(os.path.join(out_dir, "data.txt"), "w") as f:
        f.write("\n".join(texts))
    for i, text in enumerate(texts):
        sheet = render_sheet(fonts[i % len(fonts)], text, height, width)
        with open(os.path.join(out_dir, f"{i + 1}.bmp"), "wb") as f:
            f.write(bmp24_topdown_bytes(sheet))
        if not quiet and (i + 1) % 10000 == 0:
            print(f"  {i + 1}/{num_samples} sheets")
    if len(fonts) > 1:
        with open(os.path.join(out_dir, "fonts.txt"), "w") as f:
            f.write("\n".join(str(i % len(fonts)) for i in range(num_samples)))
    names = ", ".join(os.path.basename(p) if p else "PIL default font" for p in font_paths)
    with open(os.path.join(out_dir, "dataset_metadata.txt"), "w") as f:
        f.write("AI Font Renderer Dataset (Pillow/FreeType stand-in for generate_font.ts)\n"
                "==============================\n\n"
                f"Font: {names}\nFont size: {FONT_SIZE}\nSheet dimensions: {width}x{height}\nPadding: 0px\n\n"
                "Format: Images are numbered sequentially (1.bmp, 2.bmp, etc.)\n"
                "Text labels are stored line by line in data.txt (line 1 corresponds to 1.bmp)\n")
    return texts


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--font", action="append", default=None, help="TTF path; repeat for a multi-font set")
    ap.add_argument("--samples", type=int, default=150000)
    ap.add_argument("--out", default="train_input")
    args = ap.parse_args(argv)
    generate_dataset(args.out, args.samples, args.font or ["FiraCode-Retina.ttf"])
    print(f"Dataset generation complete. Check the {args.out}/ directory.")


if __name__ == "__main__":
    main()
