"""In-situ kernel timeline of one train() call (diagnostics; not a pytest file): torch.profiler (CUPTI) records the
start / duration / stream of every kernel of the replayed CUDA graphs.

    python tests/gpu_step_timeline.py [out.json] [workload]
"""
import json
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402

import os  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/step_timeline.json"
wl = sys.argv[2] if len(sys.argv) > 2 else "cfg3"
w = dict(bench.WORKLOADS[wl])
# under torchrun (data parallel): every rank trains, rank 0 records its own timeline (NCCL kernels included)
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
t = bench.build_trainer(w, "rms_prop")
batch = bench.make_batch(w, seed=42 + rank, device=dev)
for _ in range(5):
    t.train(*batch)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

if world > 1:
    dist.barrier()
    torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    t.train(*batch)
    torch.cuda.synchronize()
if rank != 0:
    dist.barrier()
    os._exit(0)
import tempfile  # noqa: E402

tmp = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(tmp)
ev = []
for e in json.load(open(tmp))["traceEvents"]:
    if e.get("cat") == "kernel":
        ev.append(dict(name=e["name"][:80], start_us=e["ts"], dur_us=e["dur"], stream=e.get("args", {}).get("stream")))
ev.sort(key=lambda x: x["start_us"])
t0 = ev[0]["start_us"] if ev else 0
for e in ev:
    e["start_us"] -= t0
json.dump(ev, open(out, "w"))
print("kernels", len(ev), "span_us", (ev[-1]["start_us"] + ev[-1]["dur_us"]) if ev else 0)
if world > 1:
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)
