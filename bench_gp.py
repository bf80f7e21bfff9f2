"""Gradient-penalty microbenchmark (BASELINE.json config 5; SURVEY.md section 8d "GP microbench figure").

    python bench_gp.py [--batches 256,1024,4096,16384] [--genes 1000,5000,18868,20000] [--iters 20]

Per (B, G): one call = the whole gradient penalty of the vanilla critic [G -> 256 -> 256 -> 1] on fp32 real / fake
[B, G]: interpolation, dD/dx_hat, GP = mean((||g|| - 1)^2) AND gp_weight * dGP/d{W1, W2, w3} (what the reference
gets from gradient_penalty + the double backward inside disc_loss.backward(), vanilla_gan_unconditional.py:304-327,
:381), through gg_engine_gp_step. The sequence is captured in a CUDA graph and timed with CUDA events, inputs
resident in HBM, L2 flushed (512 MB write) between timed iterations.

Figures per line (rank 0 prints one JSON line per shape, then a summary line):
  dense_equiv_tflops   F_GP = 8*B*G*H + 8*B*H^2 (the necessary dense count of SURVEY 8d) / time -- labelled
                       "dense-equivalent": the Gram-matrix formulation executes fewer FLOPs (executed_tflops);
  hbm_gbs / hbm_frac   algorithmic bytes 8*B*G (read fp32 real + fake once) + 4*G*H (bf16 W1 twice) + 4*G*H
                       (fp32 dW1) over time, against MEASURED_PEAKS.json hbm_gbs -- the bound of this path;
  check                GP value against a float64 autograd evaluation (oracle) on the small shapes.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="256,1024,4096,16384")
    ap.add_argument("--genes", default="1000,5000,18868,20000")
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    import torch

    import vanilla_gan_unconditional as m
    from gemmgan_b200 import _lib

    _lib.require_device(0)
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        tf_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
    except Exception:  # noqa: BLE001
        hbm_peak, tf_peak = 6650.0, 1400.0
    dev = torch.device("cuda", 0)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    H = 256
    out = []
    for G in [int(x) for x in args.genes.split(",")]:
        torch.manual_seed(42)
        t = m.WGAN_GP_nocond(input_dims=G, latent_dims=256, vocab_sizes=[], generator_dims=[H, H, G],
                             discriminator_dims=[H, H, 1], optimizer="rms_prop")
        t.build_WGAN_GP_nocond()
        t.init_train()
        for B in [int(x) for x in args.batches.split(",")]:
            eng = t._engine(B)
            g = torch.Generator(device=dev).manual_seed(1)
            real = torch.randn(B, G, device=dev, generator=g)
            fake = torch.randn(B, G, device=dev, generator=g)
            alpha = torch.rand(B, 1, device=dev, generator=g)
            gp = torch.zeros((), device=dev)
            for _ in range(3):
                eng.gp_step(real, fake, alpha, gp)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                eng.gp_step(real, fake, alpha, gp)
            ts = []
            for i in range(args.iters):
                flush.fill_(i & 0xFF)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                graph.replay()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            ms = ts[len(ts) // 2]
            dense = 8.0 * B * G * H + 8.0 * B * H * H
            executed = 2.0 * (2 * B) * G * H + 2.0 * H * H * G * 2 + 14.0 * B * H * H
            byts = 8.0 * B * G + 4.0 * G * H + 4.0 * G * H
            line = dict(bench="gp_microbench", B=B, G=G, H=H, ms=ms, dense_equiv_tflops=dense / ms / 1e9,
                        executed_tflops=executed / ms / 1e9, hbm_gbs=byts / ms / 1e6, hbm_frac=byts / ms / 1e6 / hbm_peak,
                        tensor_frac_dense_equiv=dense / ms / 1e9 / tf_peak, gp=float(gp.item()))
            if B * G <= 1024 * 5000:   # float64 autograd check of the GP value (oracle side: plain torch on the CPU)
                disc = t.disc
                W1 = disc.discriminator[0][0].weight.detach().double().cpu()
                b1 = disc.discriminator[0][0].bias.detach().double().cpu()
                W2 = disc.discriminator[1][0].weight.detach().double().cpu()
                b2 = disc.discriminator[1][0].bias.detach().double().cpu()
                w3 = disc.final_layer.weight.detach().double().cpu()
                a = alpha.double().cpu()
                xh = (a * real.double().cpu() + (1 - a) * fake.double().cpu()).requires_grad_(True)
                o = torch.relu(torch.relu(xh @ W1.t() + b1) @ W2.t() + b2) @ w3.t()
                (gr,) = torch.autograd.grad(o.sum(), xh)
                ref = ((gr.norm(dim=1) - 1) ** 2).mean().item()
                line["gp_ref_f64"] = ref
                line["gp_rel_err"] = abs(line["gp"] - ref) / max(abs(ref), 1e-12)
            print(json.dumps(line), flush=True)
            out.append(line)
            del eng
            t._engines.clear()
    best = max(out, key=lambda r: r["hbm_frac"])
    print(json.dumps(dict(bench="gp_microbench_summary", shapes=len(out), best_hbm_frac=best["hbm_frac"],
                          best_shape=[best["B"], best["G"]], hbm_peak_gbs=hbm_peak,
                          max_gp_rel_err=max((r.get("gp_rel_err", 0.0) for r in out)))))


if __name__ == "__main__":
    main()
