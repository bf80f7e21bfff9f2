"""Benchmark of the WGAN-GP training step (BASELINE.json metric: train samples/s, critic+gen steps).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3] [--impl ours|reference]

One "step" = one WGAN_GP.train() call = 5 critic steps + 1 generator step on one batch of synthetic
input (reference: src/conditional_gan_cross_attention_with_film.py:463-477). Workload at N=1: cfg3 =
the paper model (conditional_gan_cross_attention_with_film), B=1024 per GPU, G=18868 genes, 8 patch
tokens + 1 text token — the configuration BASELINE.json quotes at 1/2/4/8 B200. Data parallel runs are
weak-scaled (B per GPU fixed); N>1 is launched with torchrun (RANK / LOCAL_RANK / WORLD_SIZE).

Prints ONE JSON line (rank 0). `value`: inputs resident in HBM; `e2e`: same metric through the public
drop-in API with pinned HOST tensors (H2D copies and the loss read-back inside the timed region).
`--impl reference` times the reference's CPU path (the oracle port: /root/reference does not exist on
the GPU box) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: variant, per-GPU batch, genes, patch tokens, text tokens   (BASELINE.json configs[0..3])
    "cfg1": dict(variant="vanilla", B=64, G=5000, P=0, T=0, desc="vanilla_gan_unconditional B=64 G=5000"),
    "cfg2": dict(variant="film", B=256, G=18868, P=256, T=1, desc="conditional_gan_film B=256 G=18868 P=256"),
    "cfg3": dict(variant="paper", B=1024, G=18868, P=8, T=1,
                 desc="conditional_gan_cross_attention_with_film (paper model) B=1024/GPU G=18868 P=8 T=1"),
    "cfg4": dict(variant="paper", B=4096, G=20000, P=64, T=32,
                 desc="paper model multi-patch multi-token B=4096/GPU G=20000 P=64 T=32"),
}
METRIC = "wgan_gp_train_samples_per_sec"
UNIT = "samples/s"
# CPU arm: cfg1 / cfg3 run the FULL workload batch (cfg3: ~6 s per train() on 16 cores); cfg2 / cfg4 (10^2..10^3 s per
# call at full batch) run a reduced batch, which `sample` and `config.cpu_sample_batch` state
CPU_SAMPLE_BATCH = {"cfg1": 64, "cfg2": 8, "cfg3": 1024, "cfg4": 16}
GP_SHAPE = dict(B=16384, G=20000, H=256)   # cfg5 headline point (0.67 TFLOP dense, 2.6 GB)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(tflops=float(p["bf16_tflops_sustained"]), hbm=float(p["hbm_gbs"]), src="measured")
    except Exception:  # noqa: BLE001
        return dict(tflops=1400.0, hbm=6650.0, src="fallback")  # B200_PROFILING.md fallback (sustained)


def make_batch(w, seed, device=None, pinned=False):
    """Synthetic tensors in the reference's dataloader layout (see gemmgan_b200/synthetic.py)."""
    import torch
    from gemmgan_b200.synthetic import synthetic_tensors

    t = synthetic_tensors(w["variant"], w["B"], w["G"], max(w["P"], 1), max(w["T"], 1), seed=seed)
    if w["variant"] == "vanilla":
        args = (t[0],)
    elif w["variant"] == "film":
        text, genes, patches, ppad = t[:4]
        args = (genes, text, patches, ppad)                   # train(gene, text, patches, pad)
    else:
        text, tpad, genes, patches, ppad = t[:5]
        args = (genes, text, tpad, patches, ppad)             # train(gene, text, text_pad, patches, pad)
    if device is not None:
        return tuple(a.to(device) for a in args)
    if pinned:
        return tuple(a.pin_memory() for a in args)
    return args


def build_trainer(w, optimizer):
    import torch

    torch.manual_seed(42)  # reference default --seed 42 (:904)
    G, Hd = w["G"], 256
    if w["variant"] == "vanilla":
        import vanilla_gan_unconditional as m
        t = m.WGAN_GP_nocond(input_dims=G, latent_dims=256, vocab_sizes=[], generator_dims=[Hd, Hd, G],
                             discriminator_dims=[Hd, Hd, 1], optimizer=optimizer)
        t.build_WGAN_GP_nocond()
    else:
        m = __import__({"paper": "conditional_gan_cross_attention_with_film", "film": "conditional_gan_film"}[w["variant"]])
        t = m.WGAN_GP(input_dims=G, latent_dims=256, embedding_dims=256, generator_dims=[Hd, Hd, G],
                      discriminator_dims=[Hd, Hd, 1], optimizer=optimizer)
        t.build_WGAN_GP()
    t.init_train()
    return t


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def cpu_reference_run(w, name, steps, warmup, optimizer):
    """The reference's CPU path (oracle port) on all host cores, bounded sample of the workload."""
    import torch
    from oracle import restated

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = min(w["B"], CPU_SAMPLE_BATCH[name])
    torch.manual_seed(42)
    o = restated.OracleWGANGP(w["variant"], w["G"], optimizer=optimizer, dropout=None)  # dropout 0.1 as shipped
    x, cond = restated.synthetic_batch(w["variant"], Bs, w["G"], max(w["P"], 1), max(w["T"], 1), seed=42)
    for _ in range(warmup):
        o.train(x, cond)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.train(x, cond)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dict(value=Bs / dt, unit=UNIT, cores=cores, threads=torch.get_num_threads(), kind="port",
                sample=f"{steps} train() call(s) of {w['variant']} at B={Bs} "
                       f"({'the full workload batch' if Bs == w['B'] else 'REDUCED from the workload batch B=' + str(w['B'])}), "
                       f"G={w['G']}, P={w['P']}, T={w['T']}, fp32, {optimizer}, dropout 0.1; samples/s = B/t",
                ms_per_step=dt * 1e3)


def eager_b200_run(w, optimizer, dev, autocast, steps=5, warmup=3):
    """Stock PyTorch eager on THIS B200: the reference's step (oracle port with stock_modules=True, i.e. the
    nn.TransformerEncoder / nn.MultiheadAttention / autograd / torch.optim calls the reference makes, reference
    :144-152, :351-477) with device-resident inputs, fp32 as the reference runs it or under torch.autocast(bf16).
    This is the 'beat this' number of SURVEY.md section 8d: the reference ships no native kernels."""
    import contextlib

    import torch
    from oracle import restated

    torch.manual_seed(42)
    o = restated.OracleWGANGP(w["variant"], w["G"], optimizer=optimizer, dropout=None, device=str(dev),
                              stock_modules=True)
    x, cond = restated.synthetic_batch(w["variant"], w["B"], w["G"], max(w["P"], 1), max(w["T"], 1), seed=42,
                                       device=str(dev))
    ctx = (lambda: torch.autocast(device_type="cuda", dtype=torch.bfloat16)) if autocast else contextlib.nullcontext
    for _ in range(warmup):
        with ctx():
            o.train(x, cond)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        with ctx():
            o.train(x, cond)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del o
    torch.cuda.empty_cache()
    return dict(value=w["B"] / (ms * 1e-3), unit=UNIT, ms_per_step=ms, steps=steps, warmup=warmup,
                precision="autocast-bf16" if autocast else "fp32 (TF32 off, torch default)",
                how="oracle port with torch's stock modules (cuBLASLt + SDPA + autograd + torch.optim), device-resident "
                    "inputs, CUDA events, includes the per-step .item() syncs the reference makes")


def gp_microbench(dev, pk, iters=10):
    """BASELINE.json config 5, the 'GP-kernel TFLOP/s' half of the metric, at its headline point: the whole gradient
    penalty of the unconditional critic (value + gp_weight * dGP/d{W1, W2, w3}) through gg_engine_gp_step, CUDA-graph
    replay, CUDA events, L2 flushed between iterations. The full sweep is bench_gp.py."""
    import torch

    import vanilla_gan_unconditional as m

    B, G, H = GP_SHAPE["B"], GP_SHAPE["G"], GP_SHAPE["H"]
    torch.manual_seed(42)
    t = m.WGAN_GP_nocond(input_dims=G, latent_dims=256, vocab_sizes=[], generator_dims=[H, H, G],
                         discriminator_dims=[H, H, 1], optimizer="rms_prop")
    t.build_WGAN_GP_nocond()
    t.init_train()
    eng = t._engine(B)
    g = torch.Generator(device=dev).manual_seed(1)
    real = torch.randn(B, G, device=dev, generator=g)
    fake = torch.randn(B, G, device=dev, generator=g)
    alpha = torch.rand(B, 1, device=dev, generator=g)
    gp = torch.zeros((), device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        eng.gp_step(real, fake, alpha, gp)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        eng.gp_step(real, fake, alpha, gp)
    ts = []
    for i in range(iters):
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
    out = dict(workload=f"cfg5: gradient penalty of the unconditional critic, B={B} G={G} H={H}", ms=ms,
               dense_equivalent_tflops=dense / ms / 1e9, executed_tflops=executed / ms / 1e9,
               hbm_gbs=byts / ms / 1e6, hbm_frac=byts / ms / 1e6 / pk["hbm"],
               tensor_frac_dense_equivalent=dense / ms / 1e9 / pk["tflops"], gp=float(gp.item()),
               note="Gram-matrix formulation: executes fewer FLOPs than the dense count 8BGH + 8BH^2 (SURVEY 8d), "
                    "so the figure that bounds it is HBM: 8BG bytes of fp32 real + fake read once")
    del eng, graph, real, fake, flush
    t._engines.clear()
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS) + ["cfg5"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--optimizer", default="rms_prop", choices=["rms_prop", "adam", "adamw"])
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch (diagnostics only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the Adam leg, the stock-PyTorch-eager-on-B200 legs and the GP microbenchmark")
    ap.add_argument("--no-prefetch", action="store_true", help="e2e leg without the next-batch H2D prefetch")
    ap.add_argument("--gemm-csv", default="", help="write per-launch GEMM shapes / durations of one train() here")
    args = ap.parse_args()
    if args.workload == "cfg5":     # the GP-kernel half of the metric as its own line
        import torch
        from gemmgan_b200 import _lib

        if int(os.environ.get("RANK", "0")) != 0:
            return
        torch.cuda.set_device(0)
        _lib.require_device(0)
        pk = peaks()
        r = gp_microbench(torch.device("cuda", 0), pk, iters=max(args.steps, 5))
        print(json.dumps(dict(metric="gp_kernel_dense_equivalent_tflops", value=r["dense_equivalent_tflops"],
                              unit="TFLOP/s (dense-equivalent)", n_gpus=1, steps=max(args.steps, 5), warmup=3,
                              ms_per_step=r["ms"], higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
                              data="synthetic", config=dict(workload=r["workload"], l2="flushed between timed steps"),
                              roofline=dict(bound="hbm", achieved=r["hbm_gbs"], peak=pk["hbm"], unit="GB/s",
                                            frac=r["hbm_frac"], traffic=None), gp_microbench=r)))
        return
    w = dict(WORKLOADS[args.workload])
    if args.batch:
        w["B"] = args.batch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = dict(workload=f"{args.workload}: {w['desc']}", per_gpu_batch=w["B"], genes=w["G"], patch_tokens=w["P"],
               text_tokens=w["T"], optimizer=args.optimizer, n_critic=5, dropout=0.1,
               parallelism=f"dp{world}", l2="flushed between timed steps (512 MB write)",
               cpu_sample_batch=min(w["B"], CPU_SAMPLE_BATCH[args.workload]),
               cuda_graphs=os.environ.get("GEMMGAN_CUDA_GRAPHS", "1") != "0")

    if args.impl == "reference":
        if rank != 0:
            return
        steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
        r = cpu_reference_run(w, args.workload, steps, warmup, args.optimizer)
        line = dict(impl="reference", metric=METRIC, value=r["value"], unit=UNIT, n_gpus=args.gpus, steps=steps,
                    warmup=warmup, ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="f32", data="synthetic", config=cfg,
                    cpu_baseline=dict(value=r["value"], unit=UNIT, cores=r["cores"], kind=r["kind"], sample=r["sample"]),
                    e2e=dict(value=r["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from gemmgan_b200 import _lib

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("GEMMGAN_NUMA_BIND", "1") != "0":   # pinned batch buffers on the GPU's own NUMA node
            from gemmgan_b200.ddp import bind_to_gpu_numa_node
            cfg["numa_node"] = bind_to_gpu_numa_node(local)
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    _lib.require_device(local)
    pk = peaks()

    t = build_trainer(w, args.optimizer)
    batch_dev = make_batch(w, seed=42 + rank, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`)
    sampler = ClockSampler(local)   # nvidia-smi needs ~0.3 s to produce its first line: started before the warm-up,
                                    # stopped right after the timed region (it covers both)
    for _ in range(max(args.warmup, 3)):
        t.train(*batch_dev)
    barrier()
    L.gg_launch_count(1)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.fill_(i & 0xFF)                # evict L2 between timed iterations (outside the events)
        ev[i][0].record()
        t.train(*batch_dev)
        ev[i][1].record()
    barrier()
    clocks = sampler.stop()
    launches = int(L.gg_launch_count(0))
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_per_step = tt.item() / args.steps
    value = world * w["B"] / (ms_per_step * 1e-3)
    losses = dict(d=[float(v) for v in t.d_batch_loss], g=float(t.g_batch_loss[0]))

    # ---- live GEMM roofline leg: CUDA events around every tcgen05 GEMM launch of one more train() call. The call
    # is captured into CUDA graphs like the timed steps (the event records become external event nodes), on one
    # lane so that the bracketed kernel runs alone; the graphs are dropped afterwards.
    for eng in t._engines.values():
        eng.set_lanes(False)            # also drops the captured graphs
    t.train(*batch_dev)                 # re-capture (single lane)
    torch.cuda.synchronize()
    for eng in t._engines.values():
        eng.graphs.clear()
    t.unique_graphs = True              # every step its own capture: one event pair per launch of the whole call
    L.gg_gemm_profile_begin()
    t._replay_events.clear()
    t.train(*batch_dev)                 # capture with event nodes + replay
    t.unique_graphs = False
    ms, fl, nl = C.c_double(), C.c_double(), C.c_longlong()
    _lib.check(L.gg_gemm_profile_end(C.byref(ms), C.byref(fl), C.byref(nl)))
    el_ms, el_fl, el_by, el_n = C.c_double(), C.c_double(), C.c_double(), C.c_longlong()
    _lib.check(L.gg_enc_layer_profile(C.byref(el_ms), C.byref(el_fl), C.byref(el_by), C.byref(el_n)))
    wg_ms, wg_fl, wg_by, wg_n = C.c_double(), C.c_double(), C.c_double(), C.c_longlong()
    _lib.check(L.gg_wgrad_group_profile(C.byref(wg_ms), C.byref(wg_fl), C.byref(wg_by), C.byref(wg_n)))
    if args.gemm_csv and rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(args.gemm_csv)), exist_ok=True)
        _lib.check(L.gg_gemm_profile_dump(args.gemm_csv.encode()))
    for eng in t._engines.values():
        eng.set_lanes(True)             # drops the profiling graphs
    # ---- supplement: the same launches timed from INSIDE the kernels (%globaltimer at body start / after the
    # stores), in the real three-lane graphs. An event pair costs ~6 us around a ~6 us launch, so the event-based
    # figure above understates the in-step efficiency of the ~250 one-tile launches; this one has no such overhead
    # (but includes whatever the concurrently running lanes take away).
    in_situ = None
    try:
        cap = 4096
        stamps = torch.empty(cap, 2, dtype=torch.int64, device=dev)

        def reset():
            stamps[:, 0] = torch.iinfo(torch.int64).max
            stamps[:, 1] = 0

        for eng in t._engines.values():
            eng.graphs.clear()
        reset()
        L.gg_gemm_set_timer(C.c_void_p(stamps.data_ptr()), cap)
        t.train(*batch_dev)                  # capture (with stamp slots) + first replay
        fl_a, by_a = (C.c_double * cap)(), (C.c_double * cap)()
        n_slots = int(L.gg_gemm_timer_slots(fl_a, by_a, cap))
        torch.cuda.synchronize()
        reset()
        t.train(*batch_dev)                  # replay only
        torch.cuda.synchronize()
        L.gg_gemm_set_timer(None, 0)
        for eng in t._engines.values():
            eng.graphs.clear()
        st = stamps[:n_slots].cpu()
        dur_ns = (st[:, 1] - st[:, 0]).clamp(min=0).double()
        ok = (st[:, 1] > 0)
        tot_s = float(dur_ns[ok].sum()) * 1e-9
        fl_s = sum(fl_a[i] for i in range(n_slots) if ok[i])
        by_s = sum(by_a[i] for i in range(n_slots) if ok[i])
        if tot_s > 0:
            in_situ = dict(launches=int(ok.sum()), gemm_ms=tot_s * 1e3, hbm_gbs=by_s / tot_s / 1e9,
                           hbm_frac=by_s / tot_s / 1e9 / pk["hbm"], tensor_tflops=fl_s / tot_s / 1e12,
                           tensor_frac=fl_s / tot_s / 1e12 / pk["tflops"],
                           how="%globaltimer inside gemm_tc_kernel (body start -> stores issued), replayed three-lane graphs")
    except Exception as exc:  # noqa: BLE001 - a diagnostic must not cost the bench line
        in_situ = dict(error=repr(exc))

    # Roofline of the dominant kernel (gemm_tc_kernel): algorithmic work of the launches of one train() over their
    # summed CUDA-event durations. The step's products are skinny (K or N = 256): summed over the launches the
    # byte roofline (operands read once + outputs written once at the measured HBM copy bandwidth) is the larger
    # of the two lower bounds, so that is the bound reported; the tensor-pipe view is given next to it.
    # device time of the profiled call itself (the six graph replays; one lane, kernels serialised by the event
    # nodes): the denominator that matches ncu's serialised launch list
    torch.cuda.synchronize()
    profiled_call_ms = sum(a.elapsed_time(b) for a, b in t._replay_events) or float("nan")
    by = float(L.gg_gemm_profile_bytes())
    secs = ms.value * 1e-3
    tf = fl.value / secs / 1e12 if secs > 0 else 0.0
    gbs = by / secs / 1e9 if secs > 0 else 0.0
    t_tensor, t_hbm = fl.value / (pk["tflops"] * 1e12), by / (pk["hbm"] * 1e9)
    traffic = None
    try:  # DRAM bytes per launch of the same launches under ncu (profiles/, committed with the launch list)
        tp = os.path.join(ROOT, "profiles", "r02_gemm_traffic.json")
        with open(tp if os.path.exists(tp) else os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")) as f:
            traffic = float(json.load(f)["dram_bytes_per_launch"])
    except Exception:  # noqa: BLE001
        pass
    common = dict(in_situ=in_situ, traffic=traffic, kernel="gemm_tc_kernel (tcgen05; every launch of one train(), CUDA-event pairs inside the replayed graph, one lane)",
                  gemm_launches_per_step=int(nl.value), gemm_ms_per_step=ms.value,
                  gemm_share_of_step=ms.value / ms_per_step, gemm_ms_over_profiled_call_ms=ms.value / profiled_call_ms,
                  profiled_call_ms=profiled_call_ms, flops_per_step=fl.value, bytes_per_step=by,
                  algorithmic_bytes_per_launch=by / max(int(nl.value), 1), tensor_tflops=tf,
                  tensor_frac=tf / pk["tflops"], hbm_gbs=gbs, hbm_frac=gbs / pk["hbm"],
                  peak_source=f"{'hbm_gbs' if t_hbm >= t_tensor else 'bf16_tflops_sustained'} of {pk['src']} MEASURED_PEAKS.json")
    if t_hbm >= t_tensor:
        roofline = dict(bound="hbm", achieved=gbs, peak=pk["hbm"], unit="GB/s", frac=gbs / pk["hbm"], **common)
    else:
        roofline = dict(bound="tensor", achieved=tf, peak=pk["tflops"], unit="TFLOP/s", frac=tf / pk["tflops"], **common)
    # the fused encoder-layer kernel (enc_layer.cu: in-proj + attention + out-proj + LN + ffn + LN of one layer pass as one
    # tcgen05 launch) — the other tensor-core kernel of the step, same event-pair method, its own lower bounds
    if el_n.value > 0 and el_ms.value > 0:
        es = el_ms.value * 1e-3
        e_tf, e_gbs = el_fl.value / es / 1e12, el_by.value / es / 1e9
        e_t_tensor, e_t_hbm = el_fl.value / (pk["tflops"] * 1e12), el_by.value / (pk["hbm"] * 1e9)
        enc = dict(kernel="enc_layer_fwd_kernel (tcgen05 + TMEM + TMA, one launch per encoder layer pass)",
                   launches_per_step=int(el_n.value), ms_per_step=el_ms.value, flops_per_step=el_fl.value,
                   bytes_per_step=el_by.value, tensor_tflops=e_tf, tensor_frac=e_tf / pk["tflops"], hbm_gbs=e_gbs,
                   hbm_frac=e_gbs / pk["hbm"], bound="hbm" if e_t_hbm >= e_t_tensor else "tensor",
                   frac=max(e_tf / pk["tflops"], e_gbs / pk["hbm"]), share_of_profiled_call=el_ms.value / profiled_call_ms)
        roofline["enc_layer"] = enc
        # both tensor-core kernels together: the figure that describes the step's tcgen05 work as a whole
        tot_s = secs + es
        roofline["all_tcgen05"] = dict(launches=int(nl.value + el_n.value), ms=ms.value + el_ms.value,
                                       tensor_tflops=(fl.value + el_fl.value) / tot_s / 1e12,
                                       tensor_frac=(fl.value + el_fl.value) / tot_s / 1e12 / pk["tflops"],
                                       hbm_gbs=(by + el_by.value) / tot_s / 1e9,
                                       hbm_frac=(by + el_by.value) / tot_s / 1e9 / pk["hbm"])

    # the grouped weight-gradient kernel (wgrad_group.cu: every dW = dY^T X of one backward flush as one persistent tcgen05
    # launch, K = the token rows): its operands are activations written moments earlier (L2-resident at cfg3), so the byte
    # bound below is the HBM one an out-of-cache run would meet; the tensor view is given next to it
    if wg_n.value > 0 and wg_ms.value > 0:
        ws = wg_ms.value * 1e-3
        w_tf, w_gbs = wg_fl.value / ws / 1e12, wg_by.value / ws / 1e9
        roofline["wgrad_group"] = dict(
            kernel="wgrad_group_kernel (tcgen05, MN-major operands by 3-D TMA boxes, split-K with in-kernel reduction)",
            launches_per_step=int(wg_n.value), ms_per_step=wg_ms.value, flops_per_step=wg_fl.value,
            bytes_per_step=wg_by.value, tensor_tflops=w_tf, tensor_frac=w_tf / pk["tflops"], hbm_gbs=w_gbs,
            hbm_frac=w_gbs / pk["hbm"],
            bound="hbm" if wg_by.value / (pk["hbm"] * 1e9) >= wg_fl.value / (pk["tflops"] * 1e12) else "tensor",
            frac=max(w_tf / pk["tflops"], w_gbs / pk["hbm"]), share_of_profiled_call=wg_ms.value / profiled_call_ms)

    # ---- end-to-end through the public API with pinned host tensors
    e2e = None
    if not args.no_e2e:
        batch_host = make_batch(w, seed=42 + rank, pinned=True)
        h2d = sum(a.numel() * a.element_size() for a in batch_host)
        # The training loop a user writes: train(batch_i, prefetch=batch_{i+1}) — every step copies its own inputs
        # host->device (pinned -> HBM, inside the timed region); the copy of step i+1's inputs is issued on a copy
        # stream right after step i's kernels have been enqueued, so it overlaps them (trainer.prefetch).
        pre = None if args.no_prefetch else batch_host
        for _ in range(2):
            t.train(*batch_host, prefetch=pre)
            _ = t.d_batch_loss, t.g_batch_loss
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            t.train(*batch_host, prefetch=pre)         # H2D of the whole tuple (:465-469), overlapped when prefetched
            _ = t.d_batch_loss, t.g_batch_loss         # loss read-back (device -> pinned host)
        s1.record()
        barrier()
        te = torch.tensor([s0.elapsed_time(s1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = dict(value=world * w["B"] / (te.item() / args.steps * 1e-3), unit=UNIT, h2d_bytes_per_step=int(h2d),
                   d2h_bytes_per_step=2 * 16 * 4, ms_per_step=te.item() / args.steps,
                   h2d_overlapped_with_previous_step=pre is not None)

    # ---- extras (N=1, rank 0; each guarded: a diagnostic must not cost the bench line)
    extras = {}
    if world == 1 and not args.no_extras:
        del batch_dev
        t._engines.clear()
        t = None
        torch.cuda.empty_cache()
        other = "adam" if args.optimizer != "adam" else "rms_prop"
        try:    # the same workload under the other optimizer the report asks for (north_star names Adam, the script
                # default is RMSprop): device-resident value only
            t2 = build_trainer(w, other)
            b2 = make_batch(w, seed=42, device=dev)
            for _ in range(3):
                t2.train(*b2)
            torch.cuda.synchronize()
            n2 = max(5, min(args.steps, 10))
            tot = 0.0
            for i in range(n2):
                flush.fill_(i & 0xFF)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                t2.train(*b2)
                b.record()
                torch.cuda.synchronize()
                tot += a.elapsed_time(b)
            extras["optimizers"] = {args.optimizer: dict(value=value, ms_per_step=ms_per_step),
                                    other: dict(value=w["B"] / (tot / n2 * 1e-3), ms_per_step=tot / n2, steps=n2)}
            t2._engines.clear()
            del t2, b2
            torch.cuda.empty_cache()
        except Exception as exc:  # noqa: BLE001
            extras["optimizers"] = dict(error=repr(exc))
        for key, ac in (("fp32", False), ("autocast_bf16", True)):
            try:
                r = eager_b200_run(w, args.optimizer, dev, ac)
                r["speedup_ours_over_it"] = value / r["value"]
                extras.setdefault("torch_eager_b200", {})[key] = r
            except Exception as exc:  # noqa: BLE001
                extras.setdefault("torch_eager_b200", {})[key] = dict(error=repr(exc))
        try:
            extras["gp_microbench"] = gp_microbench(dev, pk)
        except Exception as exc:  # noqa: BLE001
            extras["gp_microbench"] = dict(error=repr(exc))

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                    ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
                    data="synthetic", config=cfg, clocks=clocks, e2e=e2e, gpu_launches=launches,
                    roofline=roofline, last_losses=losses, **extras)
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(w, args.workload, 1, 1, args.optimizer)
            line["cpu_baseline"] = dict(value=r["value"], unit=UNIT, cores=r["cores"], kind=r["kind"], sample=r["sample"])
        print(json.dumps(line), flush=True)
    if world > 1:
        # the captured step graphs hold NCCL kernels: release them before the communicator goes away, and do not
        # let a stuck teardown outlive the measurement
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
