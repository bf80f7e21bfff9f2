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
CPU_SAMPLE_BATCH = {"cfg1": 64, "cfg2": 8, "cfg3": 256, "cfg4": 16}


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
                sample=f"{steps} train() call(s) of {w['variant']} at B={Bs} (full workload B={w['B']}), "
                       f"G={w['G']}, P={w['P']}, T={w['T']}, fp32, {optimizer}, dropout 0.1; samples/s = B/t",
                ms_per_step=dt * 1e3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--optimizer", default="rms_prop", choices=["rms_prop", "adam", "adamw"])
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch (diagnostics only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-prefetch", action="store_true", help="e2e leg without the next-batch H2D prefetch")
    ap.add_argument("--gemm-csv", default="", help="write per-launch GEMM shapes / durations of one train() here")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.batch:
        w["B"] = args.batch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = dict(workload=f"{args.workload}: {w['desc']}", per_gpu_batch=w["B"], genes=w["G"], patch_tokens=w["P"],
               text_tokens=w["T"], optimizer=args.optimizer, n_critic=5, dropout=0.1,
               parallelism=f"dp{world}", l2="flushed between timed steps (512 MB write)",
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
        with open(os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")) as f:
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

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                    ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
                    data="synthetic", config=cfg, clocks=clocks, e2e=e2e, gpu_launches=launches,
                    roofline=roofline, last_losses=losses)
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
