"""CPU oracle for the ai-font-renderer hot path.  *** TEST INFRASTRUCTURE, NOT PRODUCT CODE ***

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference legs may import this
module, and only as the checker (or as the timed CPU baseline). The product path
(ai_font_renderer_b200/) never imports it and has no CPU fallback.

What it is: a plain fp32 PyTorch-on-CPU restatement of the reference's algorithm for the path
model.py:129-204 (forward), model.py:269-270 (loss), model.py:309 (backward, via autograd over the
restated forward), model.py:273/310 (AdamW, restated by hand) and helpers.py:33 (quantisation),
written with explicit tensor algebra instead of nn.MultiheadAttention so that every step cites the
statement it follows, and parametrised over (vocab, max_length, sheet size) so it can also serve
the extension configs that have no reference implementation.

Parity pinning: the reference ships no tests or golden vectors, so this
oracle is pinned against outputs of the reference module itself, imported from /root/reference in
the build container by oracle/make_golden.py; the resulting vectors live in tests/golden/ and
tests/test_oracle_golden.py checks the oracle against them on every CPU test run.
Third-party arithmetic: PyTorch (requirements.txt:2, unpinned; pinned here to 2.11.0+cu128).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

STATE_KEYS = (
    "positional_encoding", "embedding.weight", "attention.in_proj_weight",
    "attention.in_proj_bias", "attention.out_proj.weight", "attention.out_proj.bias",
    "layer_norm.weight", "layer_norm.bias", "fc1.weight", "fc1.bias",
    "fc_output.weight", "fc_output.bias")


@dataclass(frozen=True)
class OracleConfig:
    """Defaults are the reference's module constants (model.py:64-66,79-81,136,148-149)."""
    vocab: int = 128
    max_length: int = 100
    embed_dim: int = 32
    num_heads: int = 4
    hidden: int = 64
    sheet_h: int = 80
    sheet_w: int = 240
    p_embed: float = 0.2
    p_attn: float = 0.2
    p_fc1: float = 0.2 + 0.05   # model.py:149 computes DROPOUT_RATE + 0.05 in Python floats

    @property
    def K(self) -> int:
        return self.max_length * self.hidden

    @property
    def P(self) -> int:
        return self.sheet_h * self.sheet_w


# --------------------------------------------------------------------------- initialisation
def init_state(cfg: OracleConfig, seed: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """Parameters drawn exactly as AttentionFontRenderer.__init__ draws them (model.py:130-152):
    same torch.nn.init calls in the same order, so that under the same torch.manual_seed the
    state equals the reference constructor's (checked bit-for-bit in make_golden.py)."""
    if seed is not None:
        torch.manual_seed(seed)
    E, L, Fh = cfg.embed_dim, cfg.max_length, cfg.hidden
    st: Dict[str, torch.Tensor] = {}

    def linear_init(out_f, in_f):  # nn.Linear.reset_parameters
        w = torch.empty(out_f, in_f)
        torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        bound = 1 / math.sqrt(in_f)
        b = torch.empty(out_f)
        torch.nn.init.uniform_(b, -bound, bound)
        return w, b

    emb = torch.empty(cfg.vocab, E)
    torch.nn.init.normal_(emb)                                   # model.py:136 (nn.Embedding)
    pos = torch.zeros(L, E)
    torch.nn.init.normal_(pos, mean=0, std=0.02)                 # model.py:140-141
    wo, _ = linear_init(E, E)                                    # MHA.out_proj, then ...
    win = torch.empty(3 * E, E)
    torch.nn.init.xavier_uniform_(win)                           # ... MHA._reset_parameters
    w1, b1 = linear_init(Fh, E)                                  # model.py:148
    wout, bout = linear_init(cfg.P, cfg.K)                       # model.py:152
    st["positional_encoding"] = pos
    st["embedding.weight"] = emb
    st["attention.in_proj_weight"] = win
    st["attention.in_proj_bias"] = torch.zeros(3 * E)
    st["attention.out_proj.weight"] = wo
    st["attention.out_proj.bias"] = torch.zeros(E)
    st["layer_norm.weight"] = torch.ones(E)
    st["layer_norm.bias"] = torch.zeros(E)
    st["fc1.weight"] = w1
    st["fc1.bias"] = b1
    st["fc_output.weight"] = wout
    st["fc_output.bias"] = bout
    return st


# --------------------------------------------------------------------------- forward
def _dropout(x: torch.Tensor, keep: Optional[torch.Tensor], p: float) -> torch.Tensor:
    """torch's dropout arithmetic: x * (bernoulli_mask / (1 - p))."""
    if keep is None:
        return x
    noise = keep.to(x.dtype).reshape(x.shape)
    noise = noise / (1 - p)
    return x * noise


def features(state, tokens: torch.Tensor, cfg: OracleConfig, masks=None, font_ids=None) -> torch.Tensor:
    """model.py:160-193 -> [B, max_length*hidden] (the A operand of fc_output).
    font_ids (config 3, an extension of the reference): int [B]; state["font_embedding.weight"][font]
    is added to every token embedding of the sample before the embedding dropout."""
    L, E, H, Fh = cfg.max_length, cfg.embed_dim, cfg.num_heads, cfg.hidden
    dh = E // H
    x = tokens[:, : min(tokens.shape[1], L)]                      # model.py:163-164
    B, S = x.shape
    m = masks or {}
    e = state["embedding.weight"][x]                              # model.py:167
    if font_ids is not None:
        e = e + state["font_embedding.weight"][font_ids.long()].unsqueeze(1)
    e = _dropout(e, m.get("embed"), cfg.p_embed)                  # model.py:168 (before positions)
    e = e + state["positional_encoding"][:S].unsqueeze(0)         # model.py:171-172
    # nn.MultiheadAttention -> F.multi_head_attention_forward (need_weights branch)
    qkv = F.linear(e, state["attention.in_proj_weight"], state["attention.in_proj_bias"])
    q, k, v = qkv.split(E, dim=-1)
    q = q.reshape(B, S, H, dh).transpose(1, 2)                    # [B,H,S,dh], head h = channels 8h..8h+7
    k = k.reshape(B, S, H, dh).transpose(1, 2)
    v = v.reshape(B, S, H, dh).transpose(1, 2)
    q = q * math.sqrt(1.0 / float(dh))
    p = torch.softmax(q @ k.transpose(-1, -2), dim=-1)            # all S keys, no padding mask
    p = _dropout(p, m.get("attn"), cfg.p_attn)                    # attention dropout on probabilities
    ctx = (p @ v).transpose(1, 2).reshape(B, S, E)
    a = F.linear(ctx, state["attention.out_proj.weight"], state["attention.out_proj.bias"])
    h = F.layer_norm(e + a, (E,), state["layer_norm.weight"], state["layer_norm.bias"], 1e-5)  # :180
    f = torch.relu(F.linear(h, state["fc1.weight"], state["fc1.bias"]))       # model.py:183
    f = _dropout(f, m.get("fc1"), cfg.p_fc1)                                  # model.py:184
    feats = f.reshape(B, S * Fh)                                              # model.py:187
    if S < L:                                                                 # model.py:190-193
        feats = torch.cat([feats, torch.zeros(B, (L - S) * Fh, dtype=feats.dtype, device=feats.device)], dim=1)
    return feats


def logits(state, tokens, cfg: OracleConfig, masks=None, font_ids=None) -> torch.Tensor:
    """fc_output before the clamp (model.py:196) -> [B, H*W]."""
    return F.linear(features(state, tokens, cfg, masks, font_ids), state["fc_output.weight"],
                    state["fc_output.bias"])


def forward(state, tokens, cfg: OracleConfig, masks=None, font_ids=None) -> torch.Tensor:
    """AttentionFontRenderer.forward (model.py:158-204) -> [B, H, W] in [0,1]."""
    z = logits(state, tokens, cfg, masks, font_ids)
    return torch.clamp(z, 0.0, 1.0).view(-1, cfg.sheet_h, cfg.sheet_w)        # model.py:199-202


def quantise_u8(sheet: torch.Tensor) -> np.ndarray:
    """helpers.py:33: (arr * 255).astype(np.uint8) -- truncation."""
    return (sheet.detach().cpu().numpy() * 255).astype(np.uint8)


def targets_to_f32(targets_u8) -> torch.Tensor:
    """helpers.py:121: np.array(img, float32) / 255.0."""
    arr = np.asarray(targets_u8, dtype=np.float32) / 255.0
    return torch.from_numpy(arr)


# --------------------------------------------------------------------------- loss / backward
class _Bf16OutputLayerLoss(torch.autograd.Function):
    """fc_output + clamp + MSE with the GEMM operands rounded to bf16 (fp32 accumulation), i.e.
    the arithmetic the B200 kernels declare: Z = bf16(A) bf16(W)^T + b;  dZ = bf16((y-t)*mask);
    dA = s*dZ bf16(W);  dW = s*dZ^T bf16(A);  db = s*sum_b dZ  with s = 2/count. Used to pin the
    CUDA implementation tightly; the fp32 path above stays the parity target."""

    @staticmethod
    def forward(ctx, feats, w, b, t, count):
        fb = feats.to(torch.bfloat16).float()
        wb = w.to(torch.bfloat16).float()
        z = fb @ wb.t() + b
        y = torch.clamp(z, 0.0, 1.0)
        d = y - t
        resid = (d * ((z >= 0) & (z <= 1))).to(torch.bfloat16).float()
        ctx.save_for_backward(fb, wb, resid)
        ctx.count = count
        ctx.mark_non_differentiable(z)
        return (d * d).sum() / count, z

    @staticmethod
    def backward(ctx, gl, _gz):
        fb, wb, resid = ctx.saved_tensors
        s = gl * (2.0 / ctx.count)
        return s * (resid @ wb), s * (resid.t() @ fb), s * resid.sum(0), None, None


def loss_and_grads(state, tokens, targets_f32, cfg: OracleConfig, masks=None,
                   loss_count: Optional[float] = None, emulate_bf16: bool = False, font_ids=None):
    """mse_loss(model(x), t) (model.py:270,304-306) and loss.backward() (model.py:309).
    loss_count overrides the mean's denominator (data-parallel shards pass global_B*H*W).
    emulate_bf16 rounds the three GEMMs' operands to bf16 like the kernels do.
    Returns (loss, grads dict, logits)."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in state.items()}
    if emulate_bf16:
        feats = features(params, tokens, cfg, masks, font_ids)
        t = targets_f32.reshape(feats.shape[0], -1)
        count = float(loss_count) if loss_count is not None else float(t.numel())
        loss, z = _Bf16OutputLayerLoss.apply(feats, params["fc_output.weight"],
                                             params["fc_output.bias"], t, count)
    else:
        z = logits(params, tokens, cfg, masks, font_ids)
        y = torch.clamp(z, 0.0, 1.0).view(-1, cfg.sheet_h, cfg.sheet_w)
        t = targets_f32.view(y.shape)
        if loss_count is None:
            loss = F.mse_loss(y, t)
        else:
            loss = ((y - t) ** 2).sum() / loss_count
    loss.backward()
    grads = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in params.items()}
    return loss.detach(), grads, z.detach()


def logits_bf16(state, tokens, cfg: OracleConfig, masks=None) -> torch.Tensor:
    """logits() with the GEMM operands rounded to bf16 (see _Bf16OutputLayerLoss)."""
    feats = features(state, tokens, cfg, masks).to(torch.bfloat16).float()
    return feats @ state["fc_output.weight"].to(torch.bfloat16).float().t() + state["fc_output.bias"]


# --------------------------------------------------------------------------- AdamW
@dataclass
class AdamWState:
    """optim.AdamW(lr=1e-3, weight_decay=5e-4, betas=(0.9, 0.99)) as built at model.py:273."""
    lr: float = 0.001
    beta1: float = 0.9
    beta2: float = 0.99
    eps: float = 1e-8
    weight_decay: float = 0.0005
    step: int = 0
    exp_avg: Dict[str, torch.Tensor] = field(default_factory=dict)
    exp_avg_sq: Dict[str, torch.Tensor] = field(default_factory=dict)


def adamw_step(state, grads, opt: AdamWState) -> None:
    """One optimizer.step() (model.py:310), torch's single-tensor AdamW arithmetic, in place."""
    opt.step += 1
    bc1 = 1 - opt.beta1 ** opt.step
    bc2 = 1 - opt.beta2 ** opt.step
    step_size = opt.lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    for k, p in state.items():
        g = grads[k]
        if k not in opt.exp_avg:
            opt.exp_avg[k] = torch.zeros_like(p)
            opt.exp_avg_sq[k] = torch.zeros_like(p)
        m, v = opt.exp_avg[k], opt.exp_avg_sq[k]
        p.mul_(1 - opt.lr * opt.weight_decay)
        m.lerp_(g, 1 - opt.beta1)
        v.mul_(opt.beta2).addcmul_(g, g, value=1 - opt.beta2)
        denom = (v.sqrt() / bc2_sqrt).add_(opt.eps)
        p.addcdiv_(m, denom, value=-step_size)


# --------------------------------------------------------------------------- dropout masks
def _philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 on numpy uint32 arrays (same constants as the CUDA kernels)."""
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint32) for x in (c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    mask32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c0.astype(np.uint64)
        p1 = M1 * c2.astype(np.uint64)
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & mask32).astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & mask32).astype(np.uint32)
        n0 = hi1 ^ c1 ^ k0
        n2 = hi0 ^ c3 ^ k1
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = np.uint32((int(k0) + 0x9E3779B9) & 0xFFFFFFFF)
        k1 = np.uint32((int(k1) + 0xBB67AE85) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def builtin_masks(cfg: OracleConfig, B: int, S: int, seed: int, step: int, sample_offset: int = 0):
    """The keep-masks the CUDA kernels generate in dropout mode 1 (afr_sm100.h: afr_dropout,
    csrc/afr_frontend.cu): one Philox4x32-10 call, key = seed, counter = (block, site | row << 2,
    global sample g, optimizer step t), yields eight 16-bit lanes; an element keeps iff its lane is
    >= round(p * 65536).
      embedding (site 0): row 0, element i = s*E + c   -> block i // 8, lane i % 8
      attention (site 1): row h*S + s, key t           -> block t // 8, lane t % 8
      fc1       (site 2): row 0, element i = s*F + j   -> block i // 8, lane i % 8"""
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF

    def lanes(blk, c1, g):
        n = blk.shape[0]
        r = _philox4x32_10(blk, c1, np.full(n, g, np.uint32),
                           np.full(n, step & 0xFFFFFFFF, np.uint32), k0, k1)
        words = np.stack(r, axis=1)                                  # [n, 4]
        return np.stack([words & 0xFFFF, words >> 16], axis=2).reshape(n, 8)

    def flat_site(site: int, n_elem: int, p: float):
        thr = int(p * 65536.0 + 0.5)
        out = np.zeros((B, n_elem), dtype=bool)
        nblk = (n_elem + 7) // 8
        blk = np.arange(nblk, dtype=np.uint32)
        for b in range(B):
            g = np.uint32((sample_offset + b) & 0xFFFFFFFF)
            u16 = lanes(blk, np.full(nblk, site, np.uint32), g).reshape(nblk * 8)[:n_elem]
            out[b] = u16 >= thr
        return out

    def attn_site(p: float, H: int):
        thr = int(p * 65536.0 + 0.5)
        out = np.zeros((B, H * S, S), dtype=bool)
        nblk = (S + 7) // 8
        rows = np.repeat(np.arange(H * S, dtype=np.uint32), nblk)
        blk = np.tile(np.arange(nblk, dtype=np.uint32), H * S)
        c1 = (np.uint32(1) | (rows << np.uint32(2))).astype(np.uint32)
        for b in range(B):
            g = np.uint32((sample_offset + b) & 0xFFFFFFFF)
            u16 = lanes(blk, c1, g).reshape(H * S, nblk * 8)[:, :S]
            out[b] = u16 >= thr
        return out

    E, H, Fh = cfg.embed_dim, cfg.num_heads, cfg.hidden
    return {
        "embed": torch.from_numpy(flat_site(0, S * E, cfg.p_embed).reshape(B, S, E)),
        "attn": torch.from_numpy(attn_site(cfg.p_attn, H).reshape(B, H, S, S)),
        "fc1": torch.from_numpy(flat_site(2, S * Fh, cfg.p_fc1).reshape(B, S, Fh)),
    }


# --------------------------------------------------------------------------- data conventions
def lcg_text(seed: int, min_len: int = 10, max_len: int = 100) -> str:
    """generate_font.ts:164-199 in exact integer arithmetic (seed*1664525+1013904223 < 2**53, and
    floor(r*n) with r = seed/2**32 equals (seed*n) >> 32)."""
    state = seed

    def nxt(n: int) -> int:
        nonlocal state
        state = (state * 1664525 + 1013904223) % 4294967296
        return (state * n) >> 32

    length = nxt(max_len - min_len + 1) + min_len
    words = []
    remaining = length
    while remaining > 0:
        wl = min(nxt(10) + 1, remaining)
        words.append("".join(chr(65 + nxt(26)) for _ in range(wl)))
        remaining -= wl
        if remaining > 0:
            words.append(" ")
            remaining -= 1
    return "".join(words)


def dataset_strings(n: int, base_seed: int = 42):
    """generate_font.ts:203-212: sample i (0-based) uses seed i + 42."""
    return [lcg_text(i + base_seed) for i in range(n)]


def encode_strings(strings, pad_to: int) -> torch.Tensor:
    """helpers.py:57-59 / 163-177: ord() per char, right-padded with token 0."""
    out = np.zeros((len(strings), pad_to), dtype=np.int64)
    for i, s in enumerate(strings):
        codes = [ord(c) for c in s][:pad_to]
        out[i, : len(codes)] = codes
    return torch.from_numpy(out)


def synthetic_targets_u8(strings, cfg: OracleConfig, seed: int = 1234) -> np.ndarray:
    """Synthetic glyph-sheet stand-ins: white sheets with ~5 % anti-aliased ink in
    the text rows. uint8, 255 = white (helpers.py:121 convention)."""
    rng = np.random.default_rng(seed)
    n = len(strings)
    t = np.full((n, cfg.sheet_h, cfg.sheet_w), 255, dtype=np.uint8)
    levels = np.array([0, 64, 128, 192], dtype=np.uint8)
    for i, s in enumerate(strings):
        rows = min(cfg.sheet_h, max(1, math.ceil(len(s) / 33) * 14))
        ink = rng.random((rows, cfg.sheet_w)) < 0.05
        vals = levels[rng.integers(0, 4, size=(rows, cfg.sheet_w))]
        t[i, :rows][ink] = vals[ink]
    return t
