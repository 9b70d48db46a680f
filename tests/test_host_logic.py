"""CPU: host-side logic of the drop-in surface (no kernels are executed here)."""
import io
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.utils.data as tud

from conftest import REPO, load_npz
from oracle import afr_oracle as orc


def test_lcg_text_known_answers():
    from ai_font_renderer_b200.data import dataset_texts, seeded_text
    lcg = load_npz("lcg.npz")
    texts = dataset_texts(2000)
    assert texts[0] == "P JAL WZ MQWPCDYYX EOGYVE MBANVV"          # SURVEY.md section 4 KATs
    assert texts[1] == "GG U AJBHEQVVO ZFU TFI G PHRPSUL"
    assert texts[2] == "YHS IYXCTW TBALZN YHXKESJ CHFW BM"
    assert [str(s) for s in lcg["first"]] == texts[:16]
    assert np.array_equal(np.array([len(t) for t in texts]), lcg["lengths"])
    assert texts == orc.dataset_strings(2000)
    for t in texts[:200]:
        assert 10 <= len(t) <= 100 and set(t) <= set("ABCDEFGHIJKLMNOPQRSTUVWXYZ ")
        assert not t.startswith(" ") and "  " not in t
    assert seeded_text(42) == texts[0]


def test_encode_matches_reference_padding():
    from ai_font_renderer_b200.data import encode
    from ai_font_renderer_b200.render import strings_to_tokens
    t = encode(["AB C", "", "Z" * 7], 5)
    assert t.dtype == torch.int64 and t.tolist() == [[65, 66, 32, 67, 0], [0] * 5, [90] * 5]
    assert torch.equal(strings_to_tokens(["HELLO"], 8), torch.tensor([[72, 69, 76, 76, 79, 0, 0, 0]]))
    assert torch.equal(encode(["AB C"], 5), orc.encode_strings(["AB C"], 5))


def test_bmp_writer_is_byte_identical_to_pil_and_reader_round_trips(tmp_path):
    from PIL import Image
    from ai_font_renderer_b200.data import read_bmp_grey
    from ai_font_renderer_b200.render import grey_bmp_bytes
    rng = np.random.default_rng(0)
    for shape in ((80, 240), (7, 13), (1, 1), (64, 64)):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        buf = io.BytesIO()
        Image.fromarray(img).save(buf, "BMP")                      # helpers.py:36,42
        mine = grey_bmp_bytes(img)
        assert mine == buf.getvalue()
        p = tmp_path / "x.bmp"
        p.write_bytes(mine)
        assert np.array_equal(read_bmp_grey(str(p)), img)
    assert len(grey_bmp_bytes(np.zeros((80, 240), np.uint8))) == 20278


def test_batched_bmp_writer_writes_the_same_files(tmp_path):
    """render.write_bmp_files (one vectorised block + a thread pool) == one PIL save per sheet
    (helpers.py:36-42,66-68), file names string_{idx}.bmp."""
    from PIL import Image
    from ai_font_renderer_b200.render import bmp_file_block, grey_bmp_bytes, write_bmp_files
    rng = np.random.default_rng(3)
    for h, w, n in ((80, 240, 20), (9, 13, 3)):
        sheets = rng.integers(0, 256, size=(n, h, w), dtype=np.uint8)
        block = bmp_file_block(sheets)
        assert block.shape == (n, len(grey_bmp_bytes(sheets[0])))
        d = tmp_path / f"{h}x{w}"
        d.mkdir()
        write_bmp_files(sheets, str(d), first_index=5)
        for i in range(n):
            ref = d / "ref.bmp"
            Image.fromarray(sheets[i], mode="L").save(ref)
            assert (d / f"string_{5 + i}.bmp").read_bytes() == ref.read_bytes()


def test_offline_dataset_generator_follows_generate_font_ts_conventions(tmp_path):
    """fontgen (Pillow stand-in for generate_font.ts): LCG strings in data.txt, 1-based 24-bit
    top-down BMPs that helpers.load_string_dataset reads, black ink on white in the rows the
    wrapped lines occupy, multi-font sets with fonts.txt."""
    import helpers
    from ai_font_renderer_b200 import fontgen
    from ai_font_renderer_b200.data import dataset_texts, read_bmp_grey
    d = tmp_path / "train_input"
    texts = fontgen.generate_dataset(str(d), 6, [None, None], quiet=True)
    assert texts == dataset_texts(6)
    assert (d / "data.txt").read_text().split("\n") == texts
    assert (d / "fonts.txt").read_text().split() == ["0", "1", "0", "1", "0", "1"]
    raw = (d / "1.bmp").read_bytes()
    assert raw[:2] == b"BM" and int.from_bytes(raw[10:14], "little") == 54
    assert int.from_bytes(raw[18:22], "little", signed=True) == 240
    assert int.from_bytes(raw[22:26], "little", signed=True) == -80          # top-down
    assert int.from_bytes(raw[28:30], "little") == 24 and len(raw) == 54 + 720 * 80
    grey = read_bmp_grey(str(d / "1.bmp"))
    font = fontgen.load_font(None)
    n_lines = len(fontgen.wrap_text(font, texts[0], 240))
    ink_rows = np.where((grey < 128).any(axis=1))[0]
    assert grey.max() == 255 and grey.min() < 64
    assert ink_rows.min() >= 0 and ink_rows.max() <= int(n_lines * 14.4) + 4   # baseline k at (k+1)*14.4
    assert 0.005 < float((grey < 255).mean()) < 0.25
    for line in fontgen.wrap_text(font, texts[3], 240):
        assert font.getlength(line) <= 240 or " " not in line
    # the reference's contract (helpers.py:177-181): int64 tokens, float32 sheets in [0, 1]
    tokens, targets = helpers.load_string_dataset(str(d), 6).tensors
    assert targets.shape == (6, 80, 240) and targets.dtype == torch.float32
    assert float(targets.min()) >= 0.0 and float(targets.max()) == 1.0
    assert tokens.shape[0] == 6 and int(tokens[0, 0]) == ord(texts[0][0])
    # the compact form the CLI uses, and the lossless way back the trainer takes
    from ai_font_renderer_b200.data import targets_as_u8
    tok2, u8 = helpers.load_string_dataset_compact(str(d), 6)
    assert u8.dtype == torch.uint8 and torch.equal(tok2, tokens)
    assert torch.equal(targets_as_u8(targets), u8)
    assert torch.equal(u8.float() / 255.0, targets)


def test_font_control_token_encoding_and_multifont_loader(tmp_path):
    """BASELINE config 3 (extension): font id as a control token in position 0."""
    from ai_font_renderer_b200 import fontgen
    from ai_font_renderer_b200.data import encode, encode_with_font, load_multifont_dataset_u8
    from ai_font_renderer_b200.render import strings_to_tokens
    strings = ["AB C", "HELLO WORLD", ""]
    tok = encode_with_font(strings, [0, 1, 1], 12)
    assert tok.shape == (3, 12) and tok.dtype == torch.int64
    assert tok[:, 0].tolist() == [128, 129, 129]
    assert torch.equal(tok[:, 1:], encode(strings, 11))
    assert torch.equal(strings_to_tokens(strings, 12, font_ids=[0, 1, 1]), tok)
    with pytest.raises(ValueError):
        encode_with_font(strings, [0], 12)
    d = tmp_path / "mf"
    texts = fontgen.generate_dataset(str(d), 5, [None, None], quiet=True)
    tokens, targets, n_fonts = load_multifont_dataset_u8(str(d), 5)
    assert n_fonts == 2 and tokens[:, 0].tolist() == [128, 129, 128, 129, 128]
    assert int(tokens[0, 1]) == ord(texts[0][0]) and targets.shape == (5, 80, 240)


def test_reader_decodes_generate_font_ts_layout(tmp_path):
    """24-bit, BGR, top-down (negative height), rows padded to 4 bytes (generate_font.ts:6-62)."""
    from PIL import Image
    from ai_font_renderer_b200.data import read_bmp_grey
    import helpers
    rng = np.random.default_rng(1)
    h, w = 5, 7
    rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    stride = (w * 3 + 3) & ~3
    rows = np.zeros((h, stride), np.uint8)
    rows[:, : w * 3] = rgb[:, :, ::-1].reshape(h, w * 3)
    header = b"BM" + (54 + stride * h).to_bytes(4, "little") + bytes(4) + (54).to_bytes(4, "little")
    info = (40).to_bytes(4, "little") + w.to_bytes(4, "little") + (-h).to_bytes(4, "little", signed=True) \
        + (1).to_bytes(2, "little") + (24).to_bytes(2, "little") + bytes(24)
    p = tmp_path / "1.bmp"
    p.write_bytes(header + info + rows.tobytes())
    want = np.array(Image.open(str(p)).convert("L"))                 # helpers.py:118
    assert np.array_equal(read_bmp_grey(str(p)), want)
    assert np.array_equal(helpers.image_to_binary_array(str(p)), want.astype(np.float32) / 255.0)


def test_load_string_dataset_conventions_and_errors(tmp_path):
    import helpers
    from ai_font_renderer_b200.render import grey_bmp_bytes
    d = tmp_path / "train_input"
    d.mkdir()
    texts = ["AB", "HELLO WORLD", "XYZ"]
    (d / "data.txt").write_text("\n".join(texts))
    rng = np.random.default_rng(2)
    imgs = rng.integers(0, 256, (3, 8, 32), dtype=np.uint8)
    for i in range(3):
        (d / f"{i + 1}.bmp").write_bytes(grey_bmp_bytes(imgs[i]))    # 1-based file names
    ds = helpers.load_string_dataset(str(d), 3, 8, 32)
    tok, tgt = ds.tensors
    assert tok.shape == (3, 11) and tok.dtype == torch.int64 and tok[0].tolist() == [65, 66] + [0] * 9
    assert tgt.dtype == torch.float32 and np.array_equal(tgt.numpy(), imgs.astype(np.float32) / 255.0)   # helpers.py:121
    with pytest.raises(ValueError):
        helpers.load_string_dataset(str(d), 4, 8, 32)
    os.remove(d / "2.bmp")
    with pytest.raises(FileNotFoundError):
        helpers.load_string_dataset(str(d), 3, 8, 32)


def test_targets_as_u8_is_lossless_or_declines():
    from ai_font_renderer_b200.data import targets_as_u8
    u = torch.randint(0, 256, (4, 8, 32), dtype=torch.uint8)
    f = torch.from_numpy(u.numpy().astype(np.float32) / 255.0)        # helpers.py:121
    assert torch.equal(targets_as_u8(f), u)
    assert targets_as_u8(f + 1e-4) is None


def test_state_dict_layout_is_the_references():
    from ai_font_renderer_b200.renderer import AttentionFontRenderer
    from ai_font_renderer_b200 import _lib
    m = AttentionFontRenderer()
    sd = m.state_dict()
    assert tuple(sd.keys()) == _lib.STATE_DICT_KEYS == orc.STATE_KEYS
    shapes = {k: tuple(v.shape) for k, v in sd.items()}
    assert shapes == {
        "positional_encoding": (100, 32), "embedding.weight": (128, 32),
        "attention.in_proj_weight": (96, 32), "attention.in_proj_bias": (96,),
        "attention.out_proj.weight": (32, 32), "attention.out_proj.bias": (32,),
        "layer_norm.weight": (32,), "layer_norm.bias": (32,), "fc1.weight": (64, 32),
        "fc1.bias": (64,), "fc_output.weight": (19200, 6400), "fc_output.bias": (19200,)}
    assert sum(v.numel() for v in sd.values()) == 122912896
    assert all(v.dtype == torch.float32 for v in sd.values())
    assert m.max_length == 100


def test_checkpoint_round_trip(tmp_path):
    import helpers
    from ai_font_renderer_b200.renderer import AttentionFontRenderer
    cls = lambda max_length: AttentionFontRenderer(max_length=max_length, sheet_height=8, sheet_width=32)
    torch.manual_seed(5)
    m = cls(12)
    path = str(tmp_path / helpers.MODEL_FILENAME)
    helpers.save_model(m, path)
    loaded = helpers.load_model(cls, 12, filename=path)
    assert not loaded.training
    for (k, a), (_, b) in zip(m.state_dict().items(), loaded.state_dict().items()):
        assert torch.equal(a, b), k
    raw = torch.load(path)
    assert tuple(raw.keys()) == orc.STATE_KEYS


def test_cpu_inputs_fail_loudly_no_fallback():
    from ai_font_renderer_b200.renderer import AttentionFontRenderer
    m = AttentionFontRenderer(max_length=12, sheet_height=8, sheet_width=32).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros((1, 12), dtype=torch.long))


def test_batch_order_equals_reference_loaders():
    """The index-only DataLoader walks the same batches as the reference's TensorDataset loaders
    (model.py:239-266): same split, same shuffles, shared generator consumed in the same order."""
    from ai_font_renderer_b200.training import _Indices
    n, bs, seed = 1000, 64, 42
    tokens = torch.arange(n).view(n, 1)
    ds = tud.TensorDataset(tokens, torch.zeros(n, 1))
    tr, va = tud.random_split(ds, [800, 200], generator=torch.Generator().manual_seed(seed))
    g = torch.Generator(); g.manual_seed(seed)
    ref_train = tud.DataLoader(tr, batch_size=bs, shuffle=True, generator=g, num_workers=2)
    ref_val = tud.DataLoader(va, batch_size=bs, shuffle=False, generator=g, num_workers=2)
    tr2, va2 = tud.random_split(_Indices(n), [800, 200], generator=torch.Generator().manual_seed(seed))
    g2 = torch.Generator(); g2.manual_seed(seed)
    my_train = tud.DataLoader(tr2, batch_size=bs, shuffle=True, generator=g2, num_workers=0)
    my_val = tud.DataLoader(va2, batch_size=bs, shuffle=False, generator=g2, num_workers=0)
    for _epoch in range(3):
        for (x, _), idx in zip(ref_train, my_train):
            assert torch.equal(x.view(-1), idx)
        for (x, _), idx in zip(ref_val, my_val):
            assert torch.equal(x.view(-1), idx)
    assert len(my_train) == 13 and len(my_val) == 4


def test_shard_bounds_and_row_buckets():
    from ai_font_renderer_b200.training import row_buckets, shard_bounds
    for n in (1, 7, 192, 304, 1024):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    for nb in (1, 3, 8, 200):
        b = row_buckets(19200, nb)
        assert b[0][0] == 0 and b[-1][1] == 19200 and len(b) <= nb
        assert all(x[1] == y[0] for x, y in zip(b, b[1:]))
        assert all(lo % 128 == 0 and (hi - lo) % 32 == 0 for lo, hi in b)


def _declared_symbols():
    text = open(os.path.join(REPO, "include", "afr_sm100.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(afr_[a-z0-9_]+)\s*\(", text)))


def test_c_abi_library_loads_and_exports_every_declared_symbol():
    from ai_font_renderer_b200 import _lib
    lib = _lib.load()                                   # builds with nvcc if needed; no GPU calls
    declared = _declared_symbols()
    assert len(declared) >= 25
    assert sorted(_lib.SIGNATURES.keys()) == declared   # the ctypes table covers the whole header
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.lib_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (afr_[a-z0-9_]+)", out))
    assert set(declared) <= exported
    assert lib.afr_abi_version() == 1
    sass = subprocess.run(["cuobjdump", "-lelf", _lib.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in sass                            # built for B200 and nothing else


def test_cli_unknown_option_exits_1():
    res = subprocess.run([sys.executable, os.path.join(REPO, "model.py"), "--bogus"],
                         capture_output=True, text=True, cwd=REPO)
    assert res.returncode == 1
    assert "Unknown option: --bogus" in res.stdout and "Available options: --train" in res.stdout


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py --impl reference (the CPU arm of the contract): stdout is ONE JSON line with the
    metric, config, cpu_baseline and e2e objects; everything else goes to stderr."""
    import json
    import subprocess
    import sys
    from conftest import REPO
    res = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=REPO)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_glyphs_per_sec" and d["unit"] == "glyphs/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]


def test_encode_uses_ord_for_any_code_point_and_render_checks_the_vocabulary():
    """helpers.py:57 encodes with ord(): code points above 255 (CJK, the BMP render set of
    BASELINE config 5) must survive; a code point outside the embedding table is the reference's
    IndexError (model.py:167), raised on the host before anything is launched."""
    from types import SimpleNamespace
    from ai_font_renderer_b200.data import encode
    from ai_font_renderer_b200.render import check_token_range, strings_to_tokens
    tok = encode(["A\u00e9\u4e2d\uffff", "\u65e5\u672c"], 6)
    assert tok.tolist() == [[65, 0xE9, 0x4E2D, 0xFFFF, 0, 0], [0x65E5, 0x672C, 0, 0, 0, 0]]
    small = SimpleNamespace(embedding=SimpleNamespace(num_embeddings=128), max_length=6)
    big = SimpleNamespace(embedding=SimpleNamespace(num_embeddings=65536), max_length=6)
    check_token_range(big, tok)
    with pytest.raises(IndexError):
        check_token_range(small, tok)
    check_token_range(small, strings_to_tokens(["HELLO"], 6))


def test_rank_without_samples_joins_the_side_stream_before_zeroing_and_keeps_its_dropout_step(monkeypatch):
    """Data-parallel ragged last batch (global batch smaller than the world size): a rank whose
    slice is empty must (a) join the previous step's gather kernel, which may still be reading its
    peer-mapped gradient buffer, BEFORE zeroing that buffer, and (b) advance its dropout step like
    the ranks that ran a forward. Host logic only: the model is a recording stand-in."""
    from ai_font_renderer_b200 import training
    calls = []

    class P:
        def __init__(self):
            self.grad = torch.ones(3)

    class FakeModel:
        training = True
        dropout_step = 7

        def __init__(self):
            self.params = [P(), P()]

        def join_pending(self):
            calls.append(("join", [float(p.grad.sum()) for p in self.params]))

        def _ordered_params(self):
            return self.params

        def fused_forward_loss(self, *a, **k):
            calls.append(("forward",))
            self.dropout_step += 1

    monkeypatch.setattr(training, "backward_and_step",
                        lambda model, opt, buckets, world, has_samples=True, **k: calls.append(("step", has_samples)))
    tr = training.Trainer.__new__(training.Trainer)
    tr.model, tr.optimizer, tr.buckets, tr.P, tr.steps_done = FakeModel(), None, [(0, 4)], 4, 0
    tr.device = torch.device("cpu")
    tr.fonts = None
    tr.tokens = torch.zeros((8, 5), dtype=torch.int64)
    tr.targets = torch.zeros((8, 2, 2), dtype=torch.uint8)
    slot = torch.ones(())
    tr.rank, tr.world = 3, 4
    tr.train_batch(torch.arange(2), slot)                 # 2 samples, 4 ranks: rank 3 gets none
    assert calls[0] == ("join", [3.0, 3.0])               # joined while the gradients were still intact
    assert calls[1] == ("step", False)
    assert all(float(p.grad.abs().sum()) == 0 for p in tr.model.params) and float(slot) == 0.0
    assert tr.model.dropout_step == 8
    calls.clear()
    tr.rank = 0
    tr.train_batch(torch.arange(2), slot)                 # rank 0 has a sample: normal path
    assert calls == [("forward",), ("step", True)] and tr.model.dropout_step == 9
