"""N ranks x B == 1 rank x N*B (not a pytest file; run with torchrun on N GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/gpu_dp_parity.py

Every rank first trains the paper model alone on the GLOBAL batch (no process group), then the ranks train
data-parallel on their row slices with `dp_global_noise` (z / alpha drawn for the global batch and sliced).
Dropout off on both sides. Row blocks are split differently, so only the fp32 summation order differs: the
post-step weights and the last gradients must agree in relative Frobenius norm over the whole net (per-tensor
numbers are reported too, but tensors whose true gradient is zero -- e.g. the key bias of an attention
in-projection, softmax is shift-invariant -- carry rounding noise that Adam normalises to +-lr, so a per-tensor
bound is meaningless for them; the same effect bounds the whole-net weight agreement at ~lr per noisy
coordinate, hence 1e-2 on weights and 3e-2 on gradients). Prints one JSON line on rank 0.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import conditional_gan_cross_attention_with_film as m  # noqa: E402
from gemmgan_b200.synthetic import synthetic_tensors  # noqa: E402


def build(G, seed=7):
    torch.manual_seed(seed)
    t = m.WGAN_GP(input_dims=G, latent_dims=256, embedding_dims=256, generator_dims=[256, 256, G],
                  discriminator_dims=[256, 256, 1], optimizer="adam")
    t.build_WGAN_GP()
    t.init_train()
    t.dropout_p = 0.0
    return t


def run(t, batch, steps, seed=123):
    torch.manual_seed(seed)   # z / alpha stream (device generator)
    for _ in range(steps):
        t.train(*batch)
    torch.cuda.synchronize()
    out = {}
    for pre, net in (("D.", t.disc), ("G.", t.gen)):
        for k, v in net.named_parameters():
            if "patches_transformer_layer." in k:
                continue
            out[pre + k] = v.detach().float().cpu().clone()
            out["grad:" + pre + k] = v.grad.detach().float().cpu().clone()
    return out


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, G, P, T, steps = 128, 1000, 4, 2, 1
    text, tpad, genes, patches, ppad = synthetic_tensors("paper", world * B, G, P, T, seed=5, ragged=True)[:5]
    full = tuple(a.to(dev) for a in (genes, text, tpad, patches, ppad))
    single = run(build(G), full, steps)

    dist.init_process_group("nccl", device_id=dev)
    res = {}
    for overlap in (True, False):
        tag = "overlap" if overlap else "no_overlap"
        t = build(G)
        t.dp_global_noise = True
        t.dp_overlap = overlap
        mine = tuple(a[rank * B:(rank + 1) * B].contiguous() for a in full)
        dp = run(t, mine, steps)
        worst, worst_k = 0.0, ""
        num = {"w": 0.0, "g": 0.0}
        den = {"w": 0.0, "g": 0.0}
        for k, w in single.items():
            kind = "g" if k.startswith("grad:") else "w"
            d2, n2 = (dp[k] - w).norm().item() ** 2, w.norm().item() ** 2
            num[kind] += d2
            den[kind] += n2
            if kind == "w" and n2 > 0 and (d2 / n2) ** 0.5 > worst:
                worst, worst_k = (d2 / n2) ** 0.5, k
        res[tag] = dict(weights_rel_fro=(num["w"] / den["w"]) ** 0.5, grads_rel_fro=(num["g"] / den["g"]) ** 0.5,
                        worst_tensor_rel_fro=worst, worst_tensor=worst_k)
        # replicas must stay bit-identical across ranks
        flat = torch.cat([v.flatten() for v in dp.values()]).to(dev)
        ref = flat.clone()
        dist.broadcast(ref, 0)
        res[tag]["ranks_identical"] = bool(torch.equal(ref, flat))
    ok = all(v["weights_rel_fro"] < 1e-2 and v["grads_rel_fro"] < 3e-2 and v["ranks_identical"] for v in res.values())
    if rank == 0:
        print(json.dumps(dict(test="dp_parity", world=world, per_rank_batch=B, steps=steps, ok=ok, **res)))
    # the captured step graphs hold NCCL kernels: skip the communicator teardown (it can wait forever on them)
    sys.stdout.flush()
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
