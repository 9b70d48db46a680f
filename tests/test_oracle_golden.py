"""CPU: the oracle restatement against the vectors the unmodified reference produced
(oracle/make_golden.py ran /root/reference/model.py in the build container)."""
import numpy as np
import torch

from conftest import load_npz, rel_fro, state_from_npz, unpack_masks
from oracle import afr_oracle as orc

KBIAS = slice(32, 64)


def small_cfg(npz):
    v, L, h, w = (int(x) for x in npz["cfg"])
    return orc.OracleConfig(vocab=v, max_length=L, sheet_h=h, sheet_w=w)


def test_oracle_init_reproduces_reference_constructor(golden_small):
    cfg = small_cfg(golden_small)
    st = orc.init_state(cfg, seed=int(golden_small["seed"]))
    assert tuple(st.keys()) == orc.STATE_KEYS
    for k in orc.STATE_KEYS:
        assert np.array_equal(st[k].numpy(), golden_small[f"state0/{k}"]), k


def test_oracle_eval_forward_bit_exact_on_small(golden_small):
    cfg = small_cfg(golden_small)
    st = state_from_npz(golden_small, "state0")
    tokens = torch.from_numpy(golden_small["tokens"])
    y = orc.forward(st, tokens, cfg)
    assert float((y - torch.from_numpy(golden_small["y_eval"])).abs().max()) <= 1e-6
    assert np.array_equal(orc.quantise_u8(y), golden_small["q_eval"])
    short = torch.from_numpy(golden_small["short_tokens"])
    assert float((orc.forward(st, short, cfg) - torch.from_numpy(golden_small["y_short"])).abs().max()) <= 1e-6


def test_oracle_train_steps_match_reference_on_small(golden_small):
    cfg = small_cfg(golden_small)
    st = state_from_npz(golden_small, "state0")
    tokens = torch.from_numpy(golden_small["tokens"])
    targets = orc.targets_to_f32(golden_small["targets_u8"])
    opt = orc.AdamWState()
    losses = golden_small["losses"]
    for step in range(len(losses)):
        loss, grads, z = orc.loss_and_grads(st, tokens, targets, cfg, unpack_masks(golden_small, step))
        assert abs(float(loss) - losses[step]) <= 1e-6 * losses[step]
        if step == 0:
            assert rel_fro(z, golden_small["z_train0"]) <= 1e-6
            for k in orc.STATE_KEYS:
                assert rel_fro(grads[k], golden_small[f"grad0/{k}"]) < 2e-5, k
        orc.adamw_step(st, grads, opt)
    for k in orc.STATE_KEYS:
        a, b = st[k].clone(), torch.from_numpy(golden_small[f"final/{k}"].copy())
        if k == "attention.in_proj_bias":
            a[KBIAS] = 0
            b[KBIAS] = 0
        assert rel_fro(a, b) < 1e-6, k


def test_oracle_default_shape_against_reference_samples(golden_default):
    """Real shape (122.9 M parameters): weights regenerated from seed 42, results compared on the
    strided samples the fixture keeps."""
    cfg = orc.OracleConfig()
    r, c, px = (int(x) for x in golden_default["strides"])
    st = orc.init_state(cfg, seed=int(golden_default["seed"]))
    for k in orc.STATE_KEYS:
        got = st[k][::r, ::c] if k == "fc_output.weight" else st[k]
        assert np.array_equal(got.numpy(), golden_default[f"state0/{k}"]), k
        assert abs(float(st[k].double().sum()) - float(golden_default[f"sum/{k}"])) < 1e-6
    tokens = torch.from_numpy(golden_default["tokens"])
    z = orc.logits(st, tokens, cfg)
    assert rel_fro(z[:, ::px], golden_default["z_eval"]) <= 1e-6
    q = orc.quantise_u8(torch.clamp(z, 0, 1)).reshape(tokens.shape[0], -1)
    assert np.array_equal(q[:, ::px], golden_default["q_eval"])
    assert [int(x.astype(np.int64).sum()) for x in q] == [int(x) for x in golden_default["q_eval_sum"]]
    targets = orc.targets_to_f32(golden_default["targets_u8"])
    loss, grads, _ = orc.loss_and_grads(st, tokens, targets, cfg, unpack_masks(golden_default, 0))
    assert abs(float(loss) - golden_default["losses"][0]) <= 1e-6 * golden_default["losses"][0]
    for k in orc.STATE_KEYS:
        g = grads[k][::r, ::c] if k == "fc_output.weight" else grads[k]
        assert rel_fro(g, golden_default[f"grad0/{k}"]) < 2e-5, k


def test_adamw_known_answer_and_torch_equivalence():
    kat = load_npz("adamw_kat.npz")
    st = {"w": torch.from_numpy(kat["p0"].copy())}
    orc.adamw_step(st, {"w": torch.from_numpy(kat["g"].copy())}, orc.AdamWState())
    assert np.array_equal(st["w"].numpy(), kat["p1"])
    # five steps against torch.optim.AdamW itself, bit for bit
    torch.manual_seed(0)
    p = torch.nn.Parameter(torch.randn(257))
    mine = {"w": p.detach().clone()}
    opt = torch.optim.AdamW([p], lr=1e-3, weight_decay=5e-4, betas=(0.9, 0.99))
    ost = orc.AdamWState()
    for _ in range(5):
        g = torch.randn(257) * 0.01
        p.grad = g.clone()
        opt.step()
        orc.adamw_step(mine, {"w": g}, ost)
    assert torch.equal(mine["w"], p.detach())


def test_clamp_gradient_is_inclusive_at_both_ends():
    """SURVEY 2.3 row r: d clamp(z,0,1)/dz = 1 at z == 0 and z == 1 (what the loss epilogue uses)."""
    z = torch.tensor([-0.5, 0.0, 0.5, 1.0, 1.5], requires_grad=True)
    torch.clamp(z, 0.0, 1.0).sum().backward()
    assert z.grad.tolist() == [0.0, 1.0, 1.0, 1.0, 0.0]


def test_builtin_masks_are_deterministic_and_shard_invariant():
    cfg = orc.OracleConfig(max_length=12, sheet_h=8, sheet_w=32)
    a = orc.builtin_masks(cfg, 6, 12, seed=5, step=2, sample_offset=0)
    b = orc.builtin_masks(cfg, 3, 12, seed=5, step=2, sample_offset=3)
    for k in a:
        assert torch.equal(a[k][3:], b[k])
    c = orc.builtin_masks(cfg, 6, 12, seed=5, step=3, sample_offset=0)
    assert not torch.equal(a["attn"], c["attn"])
    big = orc.builtin_masks(orc.OracleConfig(), 2, 100, seed=1, step=0)
    assert abs(float(big["attn"].float().mean()) - 0.8) < 0.01
    assert abs(float(big["fc1"].float().mean()) - 0.75) < 0.02
