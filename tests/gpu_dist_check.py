"""Multi-GPU parity check on the GPU box (torchrun script, not a pytest file):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tests/gpu_dist_check.py [tiny|real2] ...

Every rank first runs the step on its own GPU alone, then with the CFG-branch x Ulysses partitioning of
flexam_b200/dist.py over NCCL, and compares the two predictions (same kernels, same accumulation order per output
element => expected bit-identical; gate 1e-3 relative L2). Also compares against the committed reference goldens when
the case matches one. Rank 0 prints one JSON line per case and appends them to gpurun_out/dist_check.json.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from flexam_b200 import dist as fdist  # noqa: E402
from flexam_b200.model import Wan2_2Transformer3DModel_FlexAM  # noqa: E402
from oracle import synth  # noqa: E402

CASES = {
    # name: (config, latent grid, golden file or None); the second grid does not divide over 4 or 8 ranks (padding)
    "tiny": ("tiny", (3, 8, 12), "tiny_tok"),
    "tiny_ragged": ("tiny", (3, 10, 18), None),
    "real2": ("real2", (5, 16, 28), "real2_tok"),
}


def rel(a, b):
    a, b = a.float(), b.float()
    return (torch.linalg.vector_norm(a - b) / (torch.linalg.vector_norm(b) + 1e-30)).item()


def build(cfg, dev):
    m = Wan2_2Transformer3DModel_FlexAM(
        model_type="ti2v", patch_size=cfg["patch_size"], text_len=cfg["text_len"], in_dim=cfg["in_dim"], dim=cfg["dim"],
        ffn_dim=cfg["ffn_dim"], freq_dim=cfg["freq_dim"], text_dim=cfg["text_dim"], out_dim=cfg["out_dim"],
        num_heads=cfg["num_heads"], num_layers=cfg["num_layers"], eps=cfg["eps"], add_ref_conv=True,
        in_dim_ref_conv=cfg["out_dim"], add_cnn_block=True, in_dim_cnn_block=cfg["in_dim_cnn"],
        out_dim_cnn_block=cfg["out_dim_cnn"], device=dev)
    m.load_state_dict(synth.state_dict_torch(cfg, dev, torch.bfloat16), strict=True)   # bit-identical to state_dict()
    return m


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    args = sys.argv[1:]
    cfg_size = 1 if "cfg1" in args else None     # "cfg1": no CFG split, all ranks sequence-parallel
    names = [a for a in args if a in CASES] or ["tiny", "tiny_ragged", "real2"]
    results, ok = [], True
    for name in names:
        cfg_name, grid, gold = CASES[name]
        cfg = synth.CONFIGS[cfg_name]
        sp = world // (cfg_size or (2 if world % 2 == 0 else 1))
        if cfg["num_heads"] % sp != 0:
            continue
        inp = synth.inputs(cfg, *grid, per_token_t=True)
        kw = dict(x=torch.from_numpy(inp["x"]).to(dev).bfloat16(), t=torch.from_numpy(inp["t"]).to(dev),
                  context=[torch.from_numpy(c).to(dev).bfloat16() for c in inp["context"]], seq_len=inp["seq_len"],
                  y=torch.from_numpy(inp["y"]).to(dev).bfloat16(),
                  full_ref=torch.from_numpy(inp["full_ref"]).to(dev).bfloat16(),
                  additional_control=torch.from_numpy(inp["additional_control"]).to(dev).bfloat16(),
                  density=torch.from_numpy(inp["density"]).to(dev))
        m = build(cfg, dev)
        single = m(**kw).clone()
        layout = fdist.setup(m, world, rank, cfg_size=cfg_size)
        m.engine()._static_key = None
        multi = m(**kw).clone()
        torch.cuda.synchronize()
        r = rel(multi, single)
        rg = None
        if gold is not None:
            g = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", gold + ".npz"))["out"]).to(dev)
            rg = rel(multi, g)
        worst = torch.tensor([r, rg or 0.0], device=dev, dtype=torch.float64)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        res = {"case": name, "world": world, "layout": layout,
               "exchange": "fused" if getattr(m.engine().par, "fused", False) else "nccl", "rel_vs_single_gpu": worst[0].item(),
               "rel_vs_reference_golden": worst[1].item() if gold else None,
               "bit_identical": bool(torch.equal(multi, single))}
        ok = ok and worst[0].item() < 1e-3 and worst[1].item() < 1e-2
        results.append(res)
        if rank == 0:
            print(json.dumps(res), flush=True)
        del m
        torch.cuda.empty_cache()
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"dist_check_n{world}.json"), "w") as f:
            json.dump(results, f, indent=1)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
