"""Times afr_train_wgrad_adamw (wgrad GEMM + AdamW epilogue) alone for a list of tile / pipeline
configurations (env knobs AFR_WA_BN / _SETS / _STAGES / _SUB read at every call) and checks
that every configuration produces bit-identical p / exp_avg / exp_avg_sq / bf16 copy.

    python tools/fused_sweep.py [--batch 1024] [--configs bn,sets,stages,sub ...]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from ai_font_renderer_b200.data import fast_synthetic_batch      # noqa: E402
from ai_font_renderer_b200.optim import FusedAdamW               # noqa: E402
from ai_font_renderer_b200.renderer import AttentionFontRenderer, _stream_ptr  # noqa: E402

DEFAULT = ["256,1,0,2", "256,2,0,1", "192,1,0,2", "128,1,0,2", "256,1,2,2"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--configs", nargs="*", default=DEFAULT)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(42)
    model = AttentionFontRenderer().to(dev).train()
    opt = FusedAdamW(model, lr=1e-3, weight_decay=5e-4, betas=(0.9, 0.99))
    tok, tgt = fast_synthetic_batch(args.batch)
    model.fused_forward_loss(tok.to(dev), tgt.to(dev))
    opt.begin_step()
    ctx, w = model._ctx, model.fc_output.weight
    P = w.shape[0]
    st = opt.state[w]
    keep = [w.data.clone(), st["exp_avg"].clone(), st["exp_avg_sq"].clone()]
    nbytes = 26 * w.numel() + 2 * args.batch * (w.shape[0] + w.shape[1])
    first = None
    for cfg in args.configs:
        bn, sets, stages, sub = (int(x) for x in cfg.split(","))
        os.environ.update(AFR_WA_BN=str(bn), AFR_WA_SETS=str(sets), AFR_WA_STAGES=str(stages),
                          AFR_WA_SUB=str(sub))
        times = []
        try:
            for rep in range(args.reps):
                for dst, src in zip((w.data, st["exp_avg"], st["exp_avg_sq"]), keep):
                    dst.copy_(src)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ctx.check(ctx.lib.afr_train_wgrad_adamw(ctx.handle, 1e-3, 0.9, 0.99, 1e-8, 5e-4, 1, 0, P,
                                                        _stream_ptr(dev)))
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
        except Exception as exc:  # configuration does not fit / invalid
            print(f"cfg bn={bn} sets={sets} stages={stages} sub={sub}: {exc}")
            continue
        shadow = ctx.workspace_tensor(3, tuple(w.shape), torch.bfloat16)
        sig = tuple(float(t.double().abs().sum()) for t in (w.data, st["exp_avg"], st["exp_avg_sq"], shadow.float()))
        if first is None:
            first = sig
        ms = min(times)
        print(f"cfg bn={bn:3d} sets={sets} stages={stages} sub={sub}: {ms:.4f} ms "
              f"(median {sorted(times)[len(times) // 2]:.4f})  {nbytes / ms / 1e6:.0f} GB/s  "
              f"{'same bits' if sig == first else 'DIFFERENT RESULT'}", flush=True)


if __name__ == "__main__":
    main()
