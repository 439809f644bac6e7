"""Host-side cost of enqueuing one denoising step (run on the GPU box; not a pytest file): builds the config-2 model,
warms up, then runs ONE step into an empty stream under cProfile and prints where the host time goes. The number that
matters at 8 GPUs, where a step is ~1,400 launches in ~44 ms.

    python tests/gpu_host_profile.py
"""
import cProfile
import io
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from flexam_b200.model import Wan2_2Transformer3DModel_FlexAM  # noqa: E402

dev = torch.device("cuda:0")
cfg = bench.real_cfg()
model = Wan2_2Transformer3DModel_FlexAM(**cfg, device=dev)
bench.init_weights(torch, model)
host = bench.make_host_inputs(torch, cfg)
d = {k: (v.to(dev) if hasattr(v, "to") else v) for k, v in host.items() if k != "context"}
d["context"] = [c.to(dev) for c in host["context"]]


def call():
    return model(x=d["x"], t=d["t"], context=d["context"], seq_len=d["seq_len"], y=d["y"], full_ref=d["full_ref"],
                 additional_control=d["additional_control"], density=d["density"])


for _ in range(2):
    call()
torch.cuda.synchronize()
for cached in (False, True):
    model.engine().cache_static = cached
    call()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    call()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"cache_static={cached}: host enqueue {1e3 * (t1 - t0):.1f} ms, step done after {1e3 * (t2 - t0):.1f} ms, "
          f"{model.engine().launches} launches")
pr = cProfile.Profile()
pr.enable()
call()
pr.disable()
torch.cuda.synchronize()
out = io.StringIO()
pstats.Stats(pr, stream=out).sort_stats("tottime").print_stats(22)
print(out.getvalue()[:6000])
