"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the REAL reference module (build container only).

    python oracle/make_golden.py            # writes tests/golden/{tiny_tok,tiny_sample,real2_tok}.npz

Each fixture stores the reference's fp32 CPU output for ``oracle.synth`` weights/inputs (which are regenerated
from names, so they are not stored), plus a few small intermediate slices used to localise a mismatch.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import flexam_oracle as O  # noqa: E402
from oracle import ref_import, synth  # noqa: E402

CASES = {
    # name: (config, (F, H, W), per-token timesteps)
    "tiny_tok": ("tiny", (3, 8, 12), True),
    "tiny_sample": ("tiny", (3, 8, 12), False),
    "real2_tok": ("real2", (5, 16, 28), True),   # BASELINE config-1 grid (560 + 112 tokens), 2 of 30 layers
}


def run_case(name: str):
    cfg_name, (F, H, W), per_tok = CASES[name]
    cfg = synth.CONFIGS[cfg_name]
    t0 = time.time()
    model = ref_import.build_reference_model(cfg).eval()
    sd = O.to_torch_sd(synth.state_dict(cfg))
    model.load_state_dict(sd, strict=True)
    inp = synth.inputs(cfg, F, H, W, per_token_t=per_tok)
    ctx = [torch.from_numpy(c) for c in inp["context"]]
    tt = {k: torch.from_numpy(inp[k]) for k in ("x", "y", "additional_control", "full_ref", "t", "density")}
    with torch.no_grad():
        out = model(x=tt["x"], t=tt["t"], context=ctx, seq_len=inp["seq_len"], y=tt["y"], full_ref=tt["full_ref"],
                    additional_control=tt["additional_control"], density=tt["density"])
        taps = {}
        mine = O.forward(sd, cfg, tt["x"], tt["t"], ctx, inp["seq_len"], tt["y"], tt["full_ref"],
                         tt["additional_control"], tt["density"], taps=taps)
    rel = ((out - mine).norm() / out.norm()).item()
    print(f"{name}: reference out {tuple(out.shape)} |max| {out.abs().max():.3f}  oracle rel-L2 {rel:.2e}  "
          f"({time.time() - t0:.1f}s)")
    assert rel < 2e-5, "oracle restatement disagrees with the reference"
    rows = slice(0, None, max(1, taps["x0"].shape[1] // 16))
    np.savez_compressed(
        os.path.join(ROOT, "tests", "golden", name + ".npz"),
        out=out.numpy().astype(np.float32),
        # oracle-side intermediates (validated end-to-end by the assert above), sample 0, every ~L/16-th token
        x0_rows=taps["x0"][0, rows].numpy(), after_self0_rows=taps["b0.after_self"][rows].numpy(),
        after_cross0_rows=taps["b0.after_cross"][rows].numpy(), x_final_rows=taps["x_final"][rows].numpy(),
        cnn_out_slice=taps["cnn_out"][0, :, 0].numpy(), ctx_rows=taps["ctx"][0, ::64].numpy(),
        meta=np.array([F, H, W, int(per_tok)], dtype=np.int64), config=np.array(cfg_name))


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for n in (sys.argv[1:] or list(CASES)):
        run_case(n)
