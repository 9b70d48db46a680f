"""BASELINE config 4 (scaled synthetic workload, widened net: 64-char strings, 64 x 64 sheets,
embed_dim 128, 8 heads, fc1 width 128) and another non-reference width: the GEMM-based front-end
(csrc/afr_wide.cu) through the C ABI against the parametrised oracle. The reference has no such
configuration (its widths are module constants, model.py:79-81,148), so this is
"restated-oracle parity" at the bf16 tolerance of north_star (2e-2): the oracle restates
model.py:158-204 for any (embed_dim, heads, hidden) and is pinned to the reference at the
reference's own widths.
"""
import ctypes as C

import pytest
import torch

from conftest import rel_fro
from oracle import afr_oracle as orc
from test_gpu_parity import dev, grads_of

pytestmark = pytest.mark.gpu

CFG4 = dict(vocab=128, max_length=64, embed_dim=128, num_heads=8, hidden=128, sheet_h=64, sheet_w=64)
ODD = dict(vocab=200, max_length=40, embed_dim=64, num_heads=4, hidden=96, sheet_h=16, sheet_w=32)


def build(cfg_kw, seed=3):
    from ai_font_renderer_b200.renderer import AttentionFontRenderer
    cfg = orc.OracleConfig(**cfg_kw)
    state = orc.init_state(cfg, seed=seed)
    m = AttentionFontRenderer(max_length=cfg.max_length, sheet_height=cfg.sheet_h, sheet_width=cfg.sheet_w,
                              vocab=cfg.vocab, embedding_dim=cfg.embed_dim, num_heads=cfg.num_heads,
                              fc1_width=cfg.hidden)
    m.load_state_dict({k: v.clone() for k, v in state.items()})
    return cfg, state, m.to(dev())


def inputs(cfg, B, seed=11, width=None):
    width = width or cfg.max_length
    strings = [s.ljust(width, "X")[:width] if i % 3 else s[:width] for i, s in enumerate(orc.dataset_strings(B, base_seed=300))]
    tokens = orc.encode_strings(strings, width)
    g = torch.Generator().manual_seed(seed)
    targets = (torch.rand((B, cfg.sheet_h, cfg.sheet_w), generator=g) < 0.8).to(torch.uint8) * 255
    ink = torch.rand((B, cfg.sheet_h, cfg.sheet_w), generator=g) < 0.3
    targets = torch.where(ink, (torch.randint(0, 4, targets.shape, generator=g) * 64).to(torch.uint8), targets)
    return tokens, targets


def kbias(cfg):
    return slice(cfg.embed_dim, 2 * cfg.embed_dim)


def assert_grads(cfg, got, want, tol, label):
    for k in orc.STATE_KEYS:
        g, w = got[k].clone(), want[k].clone()
        if k == "attention.in_proj_bias":
            g[kbias(cfg)] = 0
            w[kbias(cfg)] = 0
        e = rel_fro(g, w)
        assert e < tol, f"{label} grad {k}: rel error {e:.3e} >= {tol}"


@pytest.mark.parametrize("cfg_kw", [CFG4, ODD], ids=["config4", "odd-widths"])
def test_wide_frontend_features_and_backward_match_oracle(cfg_kw):
    """The front-end alone (debug hooks): features and the ten small gradients for an injected
    d(features), dropout off and with the built-in generator (oracle builtin_masks)."""
    cfg, state, model = build(cfg_kw)
    B = 6
    tokens, _ = inputs(cfg, B)
    S = tokens.shape[1]
    tok = tokens.to(dev())
    model.train()
    for use_drop in (False, True):
        model.dropout_seed, model.dropout_step = 77, 5
        masks = orc.builtin_masks(cfg, B, S, seed=77, step=5) if use_drop else None
        params = {k: (v.clone().requires_grad_(True) if not k.startswith("fc_output") else v) for k, v in state.items()}
        feats_ref = orc.features(params, tokens, cfg, masks)
        gen = torch.Generator().manual_seed(5)
        dfeat = torch.randn(feats_ref.shape, generator=gen) * 1e-3
        feats_ref.backward(dfeat)
        ctx = model._context(B, training=True)
        ctx.bind_grads(model._param_grads())
        drop = model.make_dropout(B, S, enabled=use_drop)
        out = torch.empty((B, cfg.K), device=dev())
        st = torch.cuda.current_stream().cuda_stream
        ctx.check(ctx.lib.afr_debug_frontend_forward(ctx.handle, tok.data_ptr(), tok.stride(0), B, S,
                                                     C.byref(drop), out.data_ptr(), st))
        # the forward GEMMs run on split-bf16 operands (~2^-16): features to 1e-4, and the same
        # elements zeroed by ReLU / dropout up to pre-activations within that error of zero
        assert rel_fro(out.cpu(), feats_ref.detach()) < 1e-4, (use_drop, rel_fro(out.cpu(), feats_ref.detach()))
        assert float(((out.cpu() == 0) != (feats_ref.detach() == 0)).float().mean()) < 1e-4
        dfd = dfeat.to(dev())
        ctx.check(ctx.lib.afr_debug_frontend_backward(ctx.handle, tok.data_ptr(), tok.stride(0), B, S,
                                                      C.byref(drop), dfd.data_ptr(), st))
        torch.cuda.synchronize()
        got = grads_of(model)
        errs = {}
        for k in orc.STATE_KEYS[:10]:
            g, w = got[k].clone(), params[k].grad.clone()
            if k == "attention.in_proj_bias":
                g[kbias(cfg)] = 0
                w[kbias(cfg)] = 0
            errs[k] = rel_fro(g, w)
        assert max(errs.values()) < 2e-2, (use_drop, errs)


@pytest.mark.parametrize("cfg_kw,B", [(CFG4, 160), (ODD, 320)], ids=["config4", "odd-widths"])
def test_wide_eval_and_train_step_match_oracle(cfg_kw, B):
    from ai_font_renderer_b200.optim import FusedAdamW
    cfg, state, model = build(cfg_kw)
    tokens, targets = inputs(cfg, B)
    z = model.eval().logits(tokens.to(dev())).cpu()
    z_ref = orc.logits(state, tokens, cfg)
    assert rel_fro(z, z_ref) < 1e-2, rel_fro(z, z_ref)
    q = model.render_u8(tokens.to(dev())).cpu().numpy()
    q_ref = orc.quantise_u8(torch.clamp(z_ref, 0, 1).view(-1, cfg.sheet_h, cfg.sheet_w))
    assert ((q >= 128) == (q_ref >= 128)).mean() >= 0.999
    model.train()
    opt = FusedAdamW(model, lr=1e-3, weight_decay=5e-4, betas=(0.9, 0.99))
    t32 = orc.targets_to_f32(targets.numpy())
    # dropout off
    loss = model.fused_train_step(tokens.to(dev()), targets.to(dev()), dropout=False)
    l_ref, g_ref, _ = orc.loss_and_grads(state, tokens, t32, cfg)
    assert abs(float(loss) - float(l_ref)) < 2e-3 * float(l_ref)
    assert_grads(cfg, grads_of(model), g_ref, 2e-2, "dropout off")
    # built-in dropout generator vs the oracle's restatement of its masks
    model.dropout_seed, model.dropout_step = 4242, 3
    masks = orc.builtin_masks(cfg, B, tokens.shape[1], seed=4242, step=3)
    loss = model.fused_train_step(tokens.to(dev()), targets.to(dev()))
    l_ref, g_ref, _ = orc.loss_and_grads(state, tokens, t32, cfg, masks)
    assert abs(float(loss) - float(l_ref)) < 2e-3 * float(l_ref)
    assert_grads(cfg, grads_of(model), g_ref, 2e-2, "philox dropout")
    # a few optimizer steps: the loss goes down and stays finite
    losses = []
    for _ in range(8):
        losses.append(float(model.fused_train_step(tokens.to(dev()), targets.to(dev()), dropout=False)))
        opt.step()
    assert all(x == x for x in losses) and losses[-1] < 0.8 * losses[0], losses


def test_wide_short_sequences_zero_tail_and_position_gradient():
    """S < max_length: features of the positions the batch never reached are zero (model.py:190-193)
    and so is their positional-encoding gradient."""
    cfg, state, model = build(CFG4)
    B, S = 5, 23
    tokens, targets = inputs(cfg, B, width=S)
    assert tokens.shape == (B, S)
    model.train()
    loss = model.fused_train_step(tokens.to(dev()), targets.to(dev()), dropout=False)
    l_ref, g_ref, _ = orc.loss_and_grads(state, tokens, orc.targets_to_f32(targets.numpy()), cfg)
    assert abs(float(loss) - float(l_ref)) < 2e-3 * float(l_ref)
    got = grads_of(model)
    assert float(got["positional_encoding"][S:].abs().max()) == 0.0
    feats = model._ctx.workspace_tensor(0, (B, cfg.K), torch.bfloat16).float().cpu()
    assert float(feats[:, S * cfg.hidden:].abs().max()) == 0.0
    assert_grads(cfg, got, g_ref, 5e-2, "short")        # B = 5: toy-batch tolerance (clamp-mask flips)


def test_reference_widths_still_take_the_fused_kernels():
    from ai_font_renderer_b200.renderer import AttentionFontRenderer
    with pytest.raises(Exception):
        m = AttentionFontRenderer(embedding_dim=48, num_heads=4).to(dev())     # head_dim 12: not built
        m.eval()(torch.zeros((1, 4), dtype=torch.int64, device=dev()))
