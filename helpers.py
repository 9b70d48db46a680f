"""Drop-in for the reference's helpers.py (same public names and conventions), implemented on the
batched B200 render path and the repo's own BMP codec.

  MODEL_FILENAME, binary_array_to_image, render_strings, save_model, load_model,
  image_to_binary_array, load_string_dataset          (helpers.py:18,20,46,76,81,107,125)
"""
import os

import numpy as np
import torch
import torch.utils.data as data

from ai_font_renderer_b200.data import load_string_dataset_u8, read_bmp_grey
from ai_font_renderer_b200.render import grey_bmp_bytes, render_strings  # noqa: F401  (re-export)

MODEL_FILENAME = "font_renderer.pth"


def binary_array_to_image(binary_array, output_path=None):
    """Grey array in [0,1] (0 = black ink, 1 = white) -> 8-bit image, optionally saved as BMP.
    Quantisation is the reference's truncating `(a * 255).astype(uint8)` (helpers.py:33)."""
    img = (np.asarray(binary_array) * 255).astype(np.uint8)
    if output_path:
        os.makedirs(os.path.dirname(output_path) or ".", exist_ok=True)
        with open(output_path, "wb") as f:
            f.write(grey_bmp_bytes(img))
    try:
        from PIL import Image
        return Image.fromarray(img)
    except ImportError:   # the hot path itself never needs PIL
        return img


def save_model(model, filename=MODEL_FILENAME):
    """state_dict -> font_renderer.pth: fp32, the reference's 12 keys (helpers.py:76-79)."""
    torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, filename)
    print(f"Model saved to {filename}")


def load_model(model_class, max_length, filename=MODEL_FILENAME, device=None):
    """helpers.py:81-105: construct, load_state_dict, move, eval()."""
    model = model_class(max_length=max_length)
    if device is None:
        device = torch.device("cpu")
    model.load_state_dict(torch.load(filename, map_location=device))
    model = model.to(device)
    model.eval()
    print(f"Model loaded from {filename}")
    return model


def image_to_binary_array(image_path):
    """BMP -> float32 [H,W] in [0,1] exactly as PIL convert('L') / 255.0 (helpers.py:107-123)."""
    return read_bmp_grey(image_path).astype(np.float32) / 255.0


def load_string_dataset(data_dir="train_input", num_samples=50000, sheet_height=80, sheet_width=240):
    """helpers.py:125-181, same contract: TensorDataset(int64 [N,Lmax], float32 [N,H,W] in [0,1]).
    (train_attention_model converts the sheets back to the lossless uint8 grey levels they came
    from -- `u8 / 255.0f` is bit-identical to this array -- before moving them to the device.)"""
    print(f"Loading {num_samples} samples from {data_dir}...")
    tokens, targets = load_string_dataset_u8(data_dir, num_samples, sheet_height, sheet_width)
    print(f"Dataset loading complete: {num_samples} samples with dimensions {sheet_height}x{sheet_width}")
    return data.TensorDataset(tokens, targets.to(torch.float32) / 255.0)


def load_string_dataset_compact(data_dir="train_input", num_samples=50000, sheet_height=80, sheet_width=240):
    """The same files as (tokens int64 [N,Lmax], uint8 [N,H,W]): a quarter of the fp32 footprint
    (2.9 GB instead of 11.5 GB for 150k samples). What `python model.py --train` itself uses."""
    print(f"Loading {num_samples} samples from {data_dir}...")
    tokens, targets = load_string_dataset_u8(data_dir, num_samples, sheet_height, sheet_width)
    print(f"Dataset loading complete: {num_samples} samples with dimensions {sheet_height}x{sheet_width}")
    return tokens, targets
