"""CPU, world_size 2 over gloo: the data-parallel host logic (sharding, global-count loss
normalisation, bucketed gradient all-reduce) with the oracle standing in for the kernels."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import REPO


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import afr_oracle as orc
    from ai_font_renderer_b200.training import row_buckets, shard_bounds
    torch.set_num_threads(1)
    cfg = orc.OracleConfig(max_length=12, sheet_h=8, sheet_w=32)
    state = orc.init_state(cfg, seed=9)
    strings = [s[:12] for s in orc.dataset_strings(10)]
    tokens = orc.encode_strings(strings, 12)
    targets = orc.targets_to_f32(orc.synthetic_targets_u8(strings, cfg, seed=3))
    gB = tokens.shape[0]
    masks = orc.builtin_masks(cfg, gB, 12, seed=77, step=1)            # keyed by GLOBAL sample index
    lo, hi = shard_bounds(gB, rank, world)
    local_masks = orc.builtin_masks(cfg, hi - lo, 12, seed=77, step=1, sample_offset=lo)
    for k in masks:
        assert torch.equal(masks[k][lo:hi], local_masks[k])
    loss, grads, _ = orc.loss_and_grads(state, tokens[lo:hi], targets[lo:hi], cfg, local_masks,
                                        loss_count=float(gB * cfg.P))
    # bucketed all-reduce of the big gradient, one flat all-reduce of the rest (training.py)
    works = []
    wg = grads["fc_output.weight"]
    for r0, r1 in row_buckets(cfg.P, 4):
        works.append(dist.all_reduce(wg[r0:r1], op=dist.ReduceOp.SUM, async_op=True))
    small_keys = [k for k in orc.STATE_KEYS if k != "fc_output.weight"]
    flat = torch.cat([grads[k].reshape(-1) for k in small_keys])
    works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True))
    for w in works:
        w.wait()
    off = 0
    for k in small_keys:
        n = grads[k].numel()
        grads[k] = flat[off:off + n].view_as(grads[k])
        off += n
    loss_t = loss.clone()
    dist.all_reduce(loss_t)
    if rank == 0:
        l_ref, g_ref, _ = orc.loss_and_grads(state, tokens, targets, cfg, masks)
        ok = abs(float(loss_t) - float(l_ref)) < 1e-6 * float(l_ref)
        worst = 0.0
        for k in orc.STATE_KEYS:
            a, b = grads[k].clone(), g_ref[k].clone()
            if k == "attention.in_proj_bias":
                a[32:64] = 0
                b[32:64] = 0
            e = float((a - b).norm() / (b.norm() + 1e-30))
            worst = max(worst, e)
        with open(os.path.join(out_dir, "result.txt"), "w") as f:
            f.write(f"{int(ok)} {worst}\n")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_data_parallel_gradients_equal_single_process(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ok, worst = open(tmp_path / "result.txt").read().split()
    assert ok == "1"
    assert float(worst) < 1e-5


def _sharded_worker(rank, world, port, out_dir):
    """training.backward_and_step's world > 1 scheme with the oracle standing in for the kernels:
    every rank ends with the same fc_output.weight as a single process stepping on the whole batch,
    although each rank ran AdamW on its own rows only."""
    import sys
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import afr_oracle as orc
    from ai_font_renderer_b200.training import owned_rows, shard_bounds
    torch.set_num_threads(1)
    cfg = orc.OracleConfig(max_length=12, sheet_h=8, sheet_w=32)
    state = orc.init_state(cfg, seed=9)
    ref_state = {k: v.clone() for k, v in state.items()}
    strings = [s[:12] for s in orc.dataset_strings(10)]
    tokens = orc.encode_strings(strings, 12)
    targets = orc.targets_to_f32(orc.synthetic_targets_u8(strings, cfg, seed=3))
    gB = tokens.shape[0]
    lo, hi = shard_bounds(gB, rank, world)
    r0, r1 = owned_rows(cfg.P, rank, world)
    opt, ref_opt = orc.AdamWState(), orc.AdamWState()
    for step in range(3):
        masks = orc.builtin_masks(cfg, gB, 12, seed=5, step=step)
        local = {k: v[lo:hi] for k, v in masks.items()}
        _, grads, _ = orc.loss_and_grads(state, tokens[lo:hi], targets[lo:hi], cfg, local,
                                         loss_count=float(gB * cfg.P))
        # reduce-scatter of dW (gloo has no reduce_scatter: all_reduce, keep the owned rows)
        wg = grads["fc_output.weight"].clone()
        dist.all_reduce(wg)
        for k in orc.STATE_KEYS:
            if k != "fc_output.weight":
                dist.all_reduce(grads[k])
        own = torch.zeros_like(wg)
        own[r0:r1] = wg[r0:r1]                       # rows of other ranks: never looked at
        grads["fc_output.weight"] = own
        before = state["fc_output.weight"].clone()
        orc.adamw_step(state, grads, opt)
        w = state["fc_output.weight"]
        w[:r0] = before[:r0]                        # only the owned rows were really swept
        w[r1:] = before[r1:]
        parts = [torch.empty_like(w[r0:r1]) for _ in range(world)]
        dist.all_gather(parts, w[r0:r1].contiguous())      # all-gather of the updated rows
        state["fc_output.weight"] = torch.cat(parts, dim=0)
        _, g_ref, _ = orc.loss_and_grads(ref_state, tokens, targets, cfg, masks)
        orc.adamw_step(ref_state, g_ref, ref_opt)
    worst = 0.0
    for k in orc.STATE_KEYS:
        a, b = state[k].clone(), ref_state[k].clone()
        if k == "attention.in_proj_bias":
            a[32:64] = 0
            b[32:64] = 0
        worst = max(worst, float((a - b).norm() / (b.norm() + 1e-30)))
    t = torch.tensor([worst])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        with open(os.path.join(out_dir, "sharded.txt"), "w") as f:
            f.write(f"{float(t[0])}\n")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_row_sharded_optimizer_equals_single_process(tmp_path):
    port = _free_port()
    mp.spawn(_sharded_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    worst = float(open(tmp_path / "sharded.txt").read())
    assert worst < 1e-5, worst
