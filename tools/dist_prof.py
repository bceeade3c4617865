"""Per-kernel timing of the slab-decomposed r2c/c2r transform (torchrun, one rank per GPU).
usage: torchrun ... tools/dist_prof.py [nx ny nz_per_rank]"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fourierflows_jl_b200 as ff
from fourierflows_jl_b200 import _lib as L

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = ff.Dist.from_torch()
a = [int(v) for v in sys.argv[1:4]] if len(sys.argv) >= 4 else [2048, 2048, 256]
shape = (a[0], a[1], a[2] * world)
T = np.float32
stream = torch.cuda.Stream()
L.call("ffb_set_stream", stream.cuda_stream)

def bar():
    dist.barrier(); torch.cuda.synchronize()

def mx(v):
    t = torch.tensor([v], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

with torch.cuda.stream(stream):
    modes = sys.argv[4].split(",") if len(sys.argv) > 4 else ["copy-engine", "peer-store", "nccl"]
    for p2p in [None if m == "nccl" else m for m in modes]:
        for wenv in (None,):
            if wenv is None: os.environ.pop("FFB_W_COLS", None)
            else: os.environ["FFB_W_COLS"] = wenv
            plan = ff.DistPlan(shape, T, comm)
            if p2p: plan.enable_p2p(p2p)
            x = ff.DevArray.zeros(T, plan.physical_shape)
            xh = ff.DevArray.zeros(ff.cxtype(T), plan.spectral_shape)
            res = {}
            for name, fn in (("rfft", lambda: plan.mul(xh, x)), ("irfft", lambda: plan.ldiv(x, xh))):
                for _ in range(3): fn()
                bar()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(8): fn()
                e1.record(stream)
                bar()
                res[name] = mx(e0.elapsed_time(e1)) / 8
                ff.prof_enable(True)
                for _ in range(2): fn()
                rep = ff.prof_report()
                ff.prof_enable(False)
                bar()
                res[name + "_k"] = [(r["name"], round(r["ms"] / r["launches"], 3)) for r in rep]
            if rank == 0:
                print(f"p2p={p2p} W_COLS={wenv}: rfft {res['rfft']:.2f} ms irfft {res['irfft']:.2f} ms", flush=True)
                print("   fwd", res["rfft_k"]); print("   inv", res["irfft_k"], flush=True)
            del x, xh, plan
            bar()
dist.destroy_process_group()
