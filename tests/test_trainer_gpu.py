"""The training DRIVER on the GPU (ai_font_renderer_b200.training.Trainer / train_attention_model,
the stand-in for model.py:209-384) against an independent re-statement of the reference's loop
driven by the CPU oracle: same split (random_split, seed 42), same shared-generator loaders
(model.py:239-266), unweighted epoch means with a ragged last batch, ReduceLROnPlateau on the
validation mean, early stopping, the files written, the checkpoint.

  * default shape, 2,240 samples, batch 1024 (train 1792 = 1024 + 768, val 448): 3 epochs;
  * small shape (12 chars, 8 x 32 sheet), 70 samples, batch 32, 60 epochs with scheduler patience 0
    and early-stopping patience 4: the LR trace and the stopping epoch must be the oracle loop's;
  * HostBatchFeeder hands out exactly the host batches, in order, reusing its two buffers.
Dropout is switched off (p = 0 on the three modules) so the two loops are comparable.
"""
import os

import numpy as np
import pytest
import torch
import torch.utils.data as tud

from conftest import rel_fro
from oracle import afr_oracle as orc
from test_gpu_parity import KBIAS, dev, make_model

pytestmark = pytest.mark.gpu


def _no_dropout(model):
    model.embedding_dropout.p = 0.0
    model.attention.dropout = 0.0
    model.dropout1.p = 0.0
    return model


def _oracle_training_loop(cfg, state, tokens, targets_f32, batch_size, epochs, lr, sched_patience,
                          early_patience, factor=0.7, min_lr=1e-6, split=0.2, seed=42):
    """model.py:232-371 with the oracle in place of the module: returns the per-epoch
    (train mean, val mean), the LR after every scheduler.step, the epoch it stopped at."""
    n = tokens.shape[0]
    val_size = int(split * n)                                                # model.py:232-233
    ds = tud.TensorDataset(tokens, targets_f32)
    train_ds, val_ds = tud.random_split(ds, [n - val_size, val_size],
                                        generator=torch.Generator().manual_seed(seed))   # :236-242
    g = torch.Generator()
    g.manual_seed(seed)                                                      # :245-246
    train_loader = tud.DataLoader(train_ds, batch_size=batch_size, shuffle=True, generator=g)   # :249-257
    val_loader = tud.DataLoader(val_ds, batch_size=batch_size, shuffle=False, generator=g)      # :258-266
    opt = orc.AdamWState(lr=lr)
    dummy = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=lr)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(dummy, mode="min", factor=factor, patience=sched_patience,
                                                       min_lr=min_lr)        # :276-278
    best, patience_counter, history, lrs, stopped_at = float("inf"), 0, [], [], None
    for epoch in range(epochs):
        total = 0.0
        for x, t in train_loader:
            opt.lr = dummy.param_groups[0]["lr"]
            loss, grads, _ = orc.loss_and_grads(state, x, t, cfg)
            orc.adamw_step(state, grads, opt)
            total += float(loss)
        vtotal = 0.0
        with torch.no_grad():
            for x, t in val_loader:
                y = orc.forward(state, x, cfg)
                vtotal += float(torch.nn.functional.mse_loss(y, t.view(y.shape)))
        avg_t, avg_v = total / len(train_loader), vtotal / len(val_loader)   # :333-334
        history.append((avg_t, avg_v))
        sched.step(avg_v)                                                    # :337
        lrs.append(dummy.param_groups[0]["lr"])
        if avg_v < best:
            best, patience_counter = avg_v, 0
        else:
            patience_counter += 1
        if patience_counter >= early_patience:                               # :362-366
            stopped_at = epoch
            break
    return history, lrs, stopped_at, best, state


def _read_kv(path):
    out = {}
    for line in open(path).read().splitlines():
        if " = " in line:
            k, v = line.split(" = ", 1)
            out[k] = v
    return out


def test_train_attention_model_default_shape_matches_oracle_loop(tmp_path):
    from ai_font_renderer_b200.data import fast_synthetic_batch
    from ai_font_renderer_b200.training import TrainConfig, train_attention_model
    cfg = orc.OracleConfig()
    n, bsz, epochs = 2240, 1024, 3
    tokens, targets_u8 = fast_synthetic_batch(n, seed=321)
    targets = targets_u8.float() / 255.0                       # the TensorDataset layout of helpers.py:177-181
    state = orc.init_state(cfg, seed=42)
    model = _no_dropout(make_model(cfg, state))
    out_dir = str(tmp_path / "train_output")
    tc = TrainConfig(output_dir=out_dir, num_epochs=epochs, test_strings=["HELLO WORLD", "AB"], render_every=1,
                     quiet=True, num_samples=n)
    train_attention_model(model, tud.TensorDataset(tokens, targets), bsz, cfg=tc, device=dev())
    tr = model._trainer
    assert len(tr.train_loader) == 2 and len(tr.val_loader) == 1          # 1024 + 768 | 448
    want_hist, want_lrs, stopped, best, ref_state = _oracle_training_loop(
        cfg, {k: v.clone() for k, v in state.items()}, tokens, targets, bsz, epochs, 1e-3, 20, 70)
    assert stopped is None and not tr.early_stopped
    for (gt, gv), (wt, wv) in zip(tr.history, want_hist):
        assert abs(gt - wt) < 2e-2 * wt and abs(gv - wv) < 2e-2 * wv, (tr.history, want_hist)
    assert tr.lr_trace == want_lrs
    # files of model.py:213-229,349-360,374-382
    kv = _read_kv(os.path.join(out_dir, "config.txt"))
    for key in ("num_epochs", "learning_rate", "batch_size", "early_stopping_patience", "validation_split",
                "weight_decay", "embedding_dim", "dropout_rate", "num_attention_heads", "max_length",
                "max_chars_per_sheet", "num_samples", "data_size", "random_seed", "sheet_height", "sheet_width"):
        assert key in kv, key
    assert kv["batch_size"] == "1024" and kv["data_size"] == str(n)
    res = _read_kv(os.path.join(out_dir, "training_results.txt"))
    assert res["final_epoch"] == str(epochs) and res["early_stopped"] == "False"
    assert abs(float(res["best_validation_loss"]) - best) < 2e-2 * best
    for e in range(epochs):
        for i in range(2):
            assert os.path.getsize(os.path.join(out_dir, f"epoch_{e}", f"string_{i}.bmp")) == 1078 + 80 * 240
    # the checkpoint: the reference's 12 keys; weights after 6 steps vs the oracle loop
    import helpers
    ckpt = str(tmp_path / "font_renderer.pth")
    helpers.save_model(model, ckpt)
    sd = torch.load(ckpt)
    assert tuple(sd.keys()) == orc.STATE_KEYS
    for k in orc.STATE_KEYS:
        a, b = sd[k].clone(), ref_state[k].clone()
        if k == "attention.in_proj_bias":
            a[KBIAS] = 0
            b[KBIAS] = 0
        tol = 1e-1 if k.endswith("bias") else 2e-2
        assert rel_fro(a, b) < tol, (k, rel_fro(a, b))


@pytest.mark.parametrize("restore_best", [False, True])
def test_scheduler_trace_and_early_stopping_match_oracle_loop(tmp_path, restore_best):
    from ai_font_renderer_b200.training import TrainConfig, train_attention_model
    cfg = orc.OracleConfig(max_length=12, sheet_h=8, sheet_w=32)
    n, bsz, epochs = 70, 32, 60
    strings = [s[:12] for s in orc.dataset_strings(n, base_seed=900)]
    tokens = orc.encode_strings(strings, 12)
    targets = orc.targets_to_f32(orc.synthetic_targets_u8(strings, cfg, seed=5))
    state = orc.init_state(cfg, seed=7)
    model = _no_dropout(make_model(cfg, state))
    tc = TrainConfig(output_dir=str(tmp_path / "out"), num_epochs=epochs, scheduler_patience=0,
                     early_stopping_patience=4, quiet=True, render_every=1000, sheet_height=8, sheet_width=32,
                     max_chars_per_sheet=12, num_samples=n, restore_best_weights=restore_best)
    train_attention_model(model, (tokens, targets), bsz, cfg=tc, device=dev())
    tr = model._trainer
    hist, lrs, stopped, best, ref_state = _oracle_training_loop(
        cfg, {k: v.clone() for k, v in state.items()}, tokens, targets, bsz, epochs, 1e-3, 0, 4)
    assert len(tr.train_loader) == 2 and len(tr.val_loader) == 1          # 56 = 32 + 24 | 14
    assert stopped is not None, "the oracle loop never stopped early: pick another workload"
    assert len(set(lrs)) > 1, "the scheduler never cut the learning rate: pick another workload"
    assert tr.early_stopped and len(tr.history) == stopped + 1, (len(tr.history), stopped)
    assert np.allclose(tr.lr_trace, lrs, rtol=1e-12), (tr.lr_trace, lrs)
    for (gt, gv), (wt, wv) in zip(tr.history, hist):
        assert abs(gt - wt) < 2e-2 * wt and abs(gv - wv) < 2e-2 * wv
    res = _read_kv(os.path.join(tc.output_dir, "training_results.txt"))
    assert res["early_stopped"] == "True" and res["final_epoch"] == str(stopped)
    # default: the reference's shallow-copy quirk (model.py:344) -> the LAST epoch's weights are
    # kept; restore_best_weights=True -> the best epoch's (different from the last epoch's)
    best_epoch = int(np.argmin([v for _, v in tr.history]))
    assert best_epoch < stopped
    final = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    last = rel_fro(final["fc_output.weight"], ref_state["fc_output.weight"])
    if restore_best:
        assert last > 1e-3, last          # not the last epoch's weights any more
        y = model.eval()(tokens.to(dev())).cpu()
        # validation loss of the restored weights == best validation loss
        val_idx = tr.val_loader.dataset.indices
        v = float(torch.nn.functional.mse_loss(y[val_idx], targets[val_idx].view(-1, 8, 32)))
        assert abs(v - tr.best_val_loss) < 1e-3 * tr.best_val_loss
    else:
        assert last < 2e-2, last


def test_host_batch_feeder_delivers_the_host_batches_in_order():
    from ai_font_renderer_b200.data import HostBatchFeeder, fast_synthetic_batch
    tok, tgt = fast_synthetic_batch(40, seed=3)
    tok, tgt = tok.pin_memory(), tgt.pin_memory()
    feeder = HostBatchFeeder(tok, tgt, 16, dev())
    assert feeder.n_batches == 2 and feeder.h2d_bytes_per_batch == 16 * (100 * 8 + 80 * 240)
    seen = set()
    for i in range(5):                       # wraps around: batch i -> host rows of batch i % 2
        x, t = feeder.get(i)
        lo = (i % 2) * 16
        assert torch.equal(x.cpu(), tok[lo:lo + 16]) and torch.equal(t.cpu(), tgt[lo:lo + 16])
        seen.add((x.data_ptr(), t.data_ptr()))
        feeder.done(i)
    assert len(seen) == 2                    # double buffered, no reallocation
    with pytest.raises(ValueError):
        HostBatchFeeder(tok.clone(), tgt.clone(), 16, dev())      # not pinned
