"""One native VAE encode (and decode) of a short clip at the real width, for launch lists under ncu:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv python profiles/r2_tools/vae_once.py encode 9
Synthetic weights come from the oracle's generators (tool code, not the product path)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from flexam_b200.vae import AutoencoderKLWan3_8  # noqa: E402
from oracle import vae_oracle as V  # noqa: E402

what, frames = sys.argv[1], int(sys.argv[2])
dev = torch.device("cuda:0")
cfg = V.VAE_CONFIGS["real"]
m = AutoencoderKLWan3_8(latent_channels=48, c_dim=cfg["enc_dim"], dec_dim=cfg["dec_dim"], device=dev)
sd = {**V.encoder_state_dict_torch(cfg, dev, torch.bfloat16), **V.state_dict_torch(cfg, dev, torch.bfloat16)}
m.load_state_dict({"model." + k: v for k, v in sd.items()}, strict=True)
if what == "encode":
    x = torch.from_numpy(V.video(cfg, frames, 512, 896)).to(dev).bfloat16()
    fn = lambda: m.encode(x).latent_dist.parameters  # noqa: E731
else:
    z = torch.from_numpy(V.latents(cfg, frames, 32, 56)).to(dev).bfloat16()
    fn = lambda: m.decode(z).sample  # noqa: E731
fn()
torch.cuda.synchronize()
torch.cuda.profiler.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
fn()
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(what, frames, "ms", e0.elapsed_time(e1), "launches", m.engine().launches)
