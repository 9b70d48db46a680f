"""`python model.py --train` / `python model.py` -- the reference's CLI (model.py:425-454) on top of
the B200 kernels.

Same surface as chenglou/ai-font-renderer's model.py: the module constants, `test_strings`, the
class `AttentionFontRenderer`, `train_attention_model`, `train_string_renderer`, the files written
(config.txt, epoch_N/string_i.bmp, training_results.txt, font_renderer.pth). The arithmetic lives in
ai_font_renderer_b200/ (hand-written sm_100a kernels behind libafr_sm100.so); this file is only the
command line. Differences a user can see:
  * a CUDA B200 is required (the reference's MPS / CPU branches, model.py:101-106, have no
    counterpart: there is no fallback path);
  * the reference pins CUDA_VISIBLE_DEVICES="3" (model.py:95); here the device is LOCAL_RANK
    (torchrun) or cuda:0;
  * optional extra flags after --train: --samples N, --epochs N, --batch N, --synthetic
    (train on synthetic sheets when train_input/ is absent; real bitmaps need bun + node-canvas;
    `python -m ai_font_renderer_b200.fontgen` writes a FreeType-rasterised stand-in), --fonts N
    (multi-font conditioning, BASELINE config 3: train_input/fonts.txt names the font of every
    sample, the font is a control token in position 0, vocabulary 128 + N, 101 positions; with
    --font-table it is a font_embedding [N, 32] table added to the token embeddings instead).
Launched under torchrun it trains data-parallel (one process per GPU, NCCL).
"""
import datetime
import os
import random
import sys

import numpy as np
import torch

from ai_font_renderer_b200.renderer import (AttentionFontRenderer, SHEET_HEIGHT, SHEET_WIDTH,
                                            MAX_CHARS_PER_SHEET, EMBEDDING_DIM, DROPOUT_RATE,
                                            NUM_ATTENTION_HEADS)
from ai_font_renderer_b200.training import TrainConfig, train_attention_model as _train
from helpers import (render_strings, save_model, load_model, load_string_dataset,  # noqa: F401
                     load_string_dataset_compact, MODEL_FILENAME)

NUM_SAMPLES = 150000
OUTPUT_DIR = "train_output_" + datetime.datetime.now().strftime("%m_%d_%H_%M_%S")

# Training hyperparameters (model.py:74-84)
NUM_EPOCHS = 10000
LEARNING_RATE = 0.001
EARLY_STOPPING_PATIENCE = 70
VALIDATION_SPLIT = 0.2
WEIGHT_DECAY = 0.0005
SCHEDULER_PATIENCE = 20
SCHEDULER_FACTOR = 0.7
MIN_LEARNING_RATE = 1e-6

SEED = 42
random.seed(SEED)
np.random.seed(SEED)
torch.manual_seed(SEED)
if torch.cuda.is_available():
    torch.cuda.manual_seed_all(SEED)

_LOCAL_RANK = int(os.environ.get("LOCAL_RANK", "0"))
if torch.cuda.is_available():
    device = torch.device("cuda", _LOCAL_RANK)
    torch.cuda.set_device(device)
    if _LOCAL_RANK == 0:
        print(f"Using CUDA device: {torch.cuda.get_device_name(device)}")
else:
    device = None   # importing is allowed (tests); running the hot path is not

if _LOCAL_RANK == 0:
    print(f"Device: {device}")

test_strings = [
    "HELLO LEANN I LOVE YOU SO MUCH I HOPE YOU HAVE A GREAT DAY",
    "TWO WORLDS ONE FAMILY TRUST YOUR HEART LET FATE DECIDE TO GUIDE THESE LIVES WE SEE",
    "A PARADISE UNTOUCHED BY MAN WITHIN THIS WORLD BLESSED WITH LOVE A SIMPLE LIFE THEY LIVE IN PEACE",
    "SOFTLY TREAD THE SAND BELOW YOUR FEET NOW TWO WORLDS ONE FAMILY TRUST YOUR HEART LET FATE",
    "BENEATH THE SHELTER OF THE TREES ONLY LOVE CAN ENTER HERE A SIMPLE LIFE THEY LIVE IN PEACE",
    "THE QUICK BROWN FOX JUMPS OVER THE LAZY DOG",
    "ABCDEFGHIJKLMNOPQRSTUVWXYZ",
    "W" * 20,
    "I" * 20,
    "ALTERNATING CASE TEST   SPACES",
    "CLAUDE IS RENDERING FONTS",
    "ZYXWVUTSRQPONMLKJIHGFEDCBA",
    "AEIOU BCDFGHJKLMNPQRSTVWXYZ",
    "EXACTLY TWENTY CHARS",
    " " * 20,
]


def _require_gpu():
    if device is None:
        raise SystemExit("model.py: no CUDA device visible -- this implementation runs on a B200 "
                         "only and has no CPU fallback")


def _config(**over):
    cfg = TrainConfig(output_dir=OUTPUT_DIR, num_epochs=NUM_EPOCHS, learning_rate=LEARNING_RATE,
                      early_stopping_patience=EARLY_STOPPING_PATIENCE,
                      validation_split=VALIDATION_SPLIT, weight_decay=WEIGHT_DECAY,
                      embedding_dim=EMBEDDING_DIM, dropout_rate=DROPOUT_RATE,
                      num_attention_heads=NUM_ATTENTION_HEADS, scheduler_patience=SCHEDULER_PATIENCE,
                      scheduler_factor=SCHEDULER_FACTOR, min_learning_rate=MIN_LEARNING_RATE,
                      seed=SEED, sheet_height=SHEET_HEIGHT, sheet_width=SHEET_WIDTH,
                      max_chars_per_sheet=MAX_CHARS_PER_SHEET, num_samples=NUM_SAMPLES,
                      test_strings=test_strings)
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg


def train_attention_model(model, dataset, batch_size, **over):
    """model.py:209."""
    return _train(model, dataset, batch_size, cfg=_config(**over), device=device)


def _flag(argv, name, default, cast=int):
    if name in argv:
        return cast(argv[argv.index(name) + 1])
    return default


def train_string_renderer(argv=()):
    """model.py:389-421."""
    _require_gpu()
    num_samples = _flag(argv, "--samples", NUM_SAMPLES)
    print("Creating sheet dataset...")
    n_fonts = _flag(argv, "--fonts", 0)
    over_fonts = {}
    if n_fonts > 0:
        pass        # the multi-font loaders below build the dataset (tokens carry the font)
    elif "--synthetic" in argv:
        from ai_font_renderer_b200.data import fast_synthetic_batch
        dataset = fast_synthetic_batch(num_samples, MAX_CHARS_PER_SHEET, SHEET_HEIGHT, SHEET_WIDTH)
    else:
        dataset = load_string_dataset_compact(data_dir="train_input", num_samples=num_samples,
                                              sheet_height=SHEET_HEIGHT, sheet_width=SHEET_WIDTH)
    if n_fonts > 0:
        from ai_font_renderer_b200.data import (dataset_texts, encode_with_font, load_multifont_dataset_u8,
                                                synthetic_sheets)
        if "--synthetic" in argv:
            texts = dataset_texts(num_samples)
            fonts = [i % n_fonts for i in range(num_samples)]
            sheets = synthetic_sheets(texts, SHEET_HEIGHT, SHEET_WIDTH)
            for f in range(1, n_fonts):          # every synthetic "font" gets its own ink level
                sheets[f::n_fonts] = 255 - (255 - sheets[f::n_fonts]) // (f + 1)
            dataset = (encode_with_font(texts, fonts, MAX_CHARS_PER_SHEET + 1), torch.from_numpy(sheets))
        else:
            tokens, targets, found = load_multifont_dataset_u8("train_input", num_samples, SHEET_HEIGHT, SHEET_WIDTH)
            if found > n_fonts:
                raise SystemExit(f"train_input/fonts.txt names {found} fonts, --fonts {n_fonts} given")
            dataset = (tokens, targets)
        over_fonts = {"test_font_ids": [i % n_fonts for i in range(len(test_strings))]}
        if "--font-table" in argv:
            # SURVEY 8d's form of the conditioning: ordinary tokens + a font_embedding [N, 32] table
            # (thirteenth checkpoint tensor) instead of the control token in position 0
            tokens_ct, sheets_ct = dataset
            over_fonts["sample_font_ids"] = (tokens_ct[:, 0] - 128).to(torch.int32)
            dataset = (tokens_ct[:, 1:].contiguous(), sheets_ct)
    print("Training attention-based sheet renderer with reduced embedding dimensions (32) and "
          "learned positional encoding...")
    if n_fonts > 0 and "--font-table" in argv:
        model = AttentionFontRenderer(max_length=dataset[0].shape[1], n_fonts=n_fonts).to(device)
    elif n_fonts > 0:
        model = AttentionFontRenderer(max_length=dataset[0].shape[1], vocab=128 + n_fonts).to(device)
    else:
        model = AttentionFontRenderer(max_length=MAX_CHARS_PER_SHEET).to(device)
    batch_size = _flag(argv, "--batch", 1024)           # model.py:408-409 (GPU batch)
    print(f"Using batch size {batch_size}")
    over = {"num_samples": num_samples, **over_fonts}
    if "--epochs" in argv:
        over["num_epochs"] = _flag(argv, "--epochs", NUM_EPOCHS)
    return train_attention_model(model, dataset, batch_size, **over)


def _init_distributed():
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=device)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        if sys.argv[1] == "--train":
            _require_gpu()
            _init_distributed()
            if _LOCAL_RANK == 0:
                os.makedirs(OUTPUT_DIR, exist_ok=True)
            model = train_string_renderer(sys.argv[2:])
            if int(os.environ.get("RANK", "0")) == 0:
                n_fonts = _flag(sys.argv[2:], "--fonts", 0)
                if n_fonts > 0:
                    # a multi-font model (vocabulary 128 + N, one more position) is not loadable by
                    # plain `python model.py`: it gets its own file name and is rendered with font ids
                    save_model(model, f"font_renderer_{n_fonts}fonts.pth")
                    render_strings(model, test_strings, output_dir=OUTPUT_DIR, sheet_height=SHEET_HEIGHT,
                                   sheet_width=SHEET_WIDTH, device=device,
                                   font_ids=[i % n_fonts for i in range(len(test_strings))])
                else:
                    save_model(model)
                    render_strings(model, test_strings, output_dir=OUTPUT_DIR, sheet_height=SHEET_HEIGHT,
                                   sheet_width=SHEET_WIDTH, device=device)
        else:
            print(f"Unknown option: {sys.argv[1]}")
            print("Available options: --train")
            sys.exit(1)
    else:
        _require_gpu()
        os.makedirs(OUTPUT_DIR, exist_ok=True)
        if os.path.exists(MODEL_FILENAME):
            model = load_model(AttentionFontRenderer, MAX_CHARS_PER_SHEET, device=device)
        else:
            print("No saved model found. Training a new model...")
            model = train_string_renderer()
            save_model(model)
        render_strings(model, test_strings, output_dir=OUTPUT_DIR, sheet_height=SHEET_HEIGHT,
                       sheet_width=SHEET_WIDTH, device=device)
