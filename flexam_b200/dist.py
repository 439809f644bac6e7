"""Multi-GPU partitioning of the denoising step on one NVLink/NVSwitch box: CFG-branch parallelism x Ulysses
sequence parallelism. This is the replacement for the reference's absent ``FlexAM/dist`` package (SURVEY.md F1;
hook sites wan_transformer3d_FlexAM.py:801-815, :919-920, :971-975, :1103-1104).

One process per GPU (torchrun). Every rank runs the unchanged sampler with the same inputs, so ``forward`` still
receives the batch-of-2 ``[uncond, cond]`` tensors (pipeline :850). Inside ``forward`` a rank keeps

  * CFG:      row ``cfg_rank`` of the batch (the two branches are independent until the combine, pipeline :926-928);
  * Ulysses:  token slice ``sp_rank`` of length Lp = ceil(L / P) — everything except self-attention is token-local.
              Around self-attention the q/k/v shards ``[Lp, H, 128]`` are exchanged head-scattered to
              ``[P*Lp, H/P, 128]`` (one all-to-all per tensor), attention runs on H/P heads over the whole sequence
              (padding tokens masked as keys through ``Lk = L``), and the output goes back through one all-to-all.

and before returning all-gathers the head output over the SP group (dim 1, as :1103-1104) and the prediction over
the CFG group (dim 0), so the caller sees the same ``[2, C, F, H, W]`` tensor as on one GPU.

Layouts: 2 GPUs = CFG2 x SP1, 4 = CFG2 x SP2, 8 = CFG2 x SP4 (24 heads -> 6 per rank).

Two implementations of the exchange around self-attention:

  * fused (GPUs, the product path): no collective call at all. The q|k RMSNorm+RoPE kernel writes every head straight
    into the exchange buffer of the rank that owns it, and the attention kernel writes every output row straight into
    the o-projection input of the rank that owns the token - plain 16-byte stores to peer-mapped memory over
    NVLink/NVSwitch (``fx_qkv_norm_rope_scatter`` / ``fx_fmha_fwd_scatter``). The buffers are torch symmetric-memory
    allocations (plumbing: allocation + peer pointers); two stream-ordered cross-rank barriers per layer order the
    peer writes against their readers. Pack kernels, four all-to-alls and the unpack kernel per layer disappear.
  * collective (``FLEXAM_SP_EXCHANGE=nccl``, and the gloo CPU tests): NCCL ``all_to_all_single`` with the native
    ``fx_swap01_bf16`` pack/unpack kernels. The functions take the permute/attention kernels as arguments so the CPU
    tests (world_size 2, gloo) drive the same partitioning logic with torch stand-ins.

The token all-gather of the head output and the CFG all-gather of the prediction (once per step) stay NCCL.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, List, Optional

import torch
import torch.distributed as dist


@dataclass
class SymmBuffer:
    tensor: torch.Tensor
    handle: object
    ptrs: List[int]


@dataclass
class Layout:
    world: int
    rank: int
    cfg_size: int
    sp_size: int

    @property
    def cfg_rank(self) -> int:
        return self.rank // self.sp_size

    @property
    def sp_rank(self) -> int:
        return self.rank % self.sp_size

    def sp_ranks(self, cfg_rank: int) -> List[int]:
        return [cfg_rank * self.sp_size + i for i in range(self.sp_size)]

    def cfg_ranks(self, sp_rank: int) -> List[int]:
        return [c * self.sp_size + sp_rank for c in range(self.cfg_size)]

    def describe(self) -> str:
        return f"cfg{self.cfg_size}xsp{self.sp_size}"


def make_layout(world: int, rank: int, cfg_size: Optional[int] = None, num_heads: int = 24) -> Layout:
    """CFG-parallel first (no per-layer traffic), then Ulysses over what is left."""
    if cfg_size is None:
        env = os.environ.get("FLEXAM_CFG_SIZE")       # e.g. 1: sequence parallelism only (SP8 with batched CFG)
        cfg_size = int(env) if env else (2 if world % 2 == 0 else 1)
    if world % cfg_size != 0:
        raise ValueError(f"world size {world} not divisible by cfg_size {cfg_size}")
    sp = world // cfg_size
    if num_heads % sp != 0:
        raise ValueError(f"{num_heads} heads do not divide over {sp} sequence-parallel ranks")
    return Layout(world, rank, cfg_size, sp)


def shard_bounds(L: int, P: int, r: int):
    """Token slice of rank r: (start, stop, Lp, L_pad); the last rank(s) may hold padding rows (:919-920)."""
    Lp = -(-L // P)
    return r * Lp, min((r + 1) * Lp, L), Lp, Lp * P


def _all_to_all(out: torch.Tensor, inp: torch.Tensor, group) -> None:
    """all_to_all_single over dim 0; gloo (CPU tests) has no all-to-all, so it is composed from all_gather there."""
    if dist.get_backend(group) == "nccl":
        dist.all_to_all_single(out, inp, group=group)
        return
    P = dist.get_world_size(group)
    me = dist.get_rank(group)
    gathered = [torch.empty_like(inp) for _ in range(P)]
    dist.all_gather(gathered, inp.contiguous(), group=group)
    chunk = inp.shape[0] // P
    for src in range(P):
        out[src * chunk:(src + 1) * chunk].copy_(gathered[src][me * chunk:(me + 1) * chunk])


class Parallel:
    """Process groups + the exchange steps used by NativeEngine.forward."""

    def __init__(self, layout: Layout, swap01: Callable, fmha: Callable):
        self.layout = layout
        self.swap01 = swap01
        self.fmha = fmha
        self.sp_group = None
        self.cfg_group = None
        # every rank must create every group, in the same order
        for c in range(layout.cfg_size):
            g = dist.new_group(layout.sp_ranks(c))
            if c == layout.cfg_rank:
                self.sp_group = g
        for s in range(layout.sp_size):
            g = dist.new_group(layout.cfg_ranks(s))
            if s == layout.sp_rank:
                self.cfg_group = g
        self._bufs = {}
        self._symm = {}
        self.fused = (os.environ.get("FLEXAM_SP_EXCHANGE", "fused") != "nccl" and layout.sp_size > 1 and
                      dist.get_backend(self.sp_group) == "nccl")

    # -- symmetric (peer-mapped) buffers for the fused exchange ---------------------------------------------------
    def symm_buffer(self, name: str, shape, dtype, device) -> "SymmBuffer":
        """A buffer every rank of the SP group allocates with the same shape, plus the device addresses of all
        ranks' copies. Collective over the SP group on first use of a (name, shape)."""
        key = (name, tuple(shape), dtype)
        b = self._symm.get(key)
        if b is None:
            import torch.distributed._symmetric_memory as symm_mem
            for k in [k for k in self._symm if k[0] == name]:
                del self._symm[k]
            t = symm_mem.empty(*shape, dtype=dtype, device=device)
            hdl = symm_mem.rendezvous(t, self.sp_group)
            b = SymmBuffer(t, hdl, [int(p) for p in hdl.buffer_ptrs])
            self._symm[key] = b
        return b

    def attention_fused(self, qkv: torch.Tensor, attn: "SymmBuffer", w_q, w_k, eps, freqs, grid, L: int, scale: float,
                        ops, timed=None) -> None:
        """qkv: [B*Lp, 3D] packed projection output of the local token slices; attn: the symmetric [B*Lp, D] input
        buffer of the o projection (filled by the peers). One scatter kernel, barrier, attention with scattered
        output, barrier."""
        lay = self.layout
        P = lay.sp_size
        D = qkv.shape[1] // 3
        B = attn.tensor.shape[0]
        Lp = attn.tensor.shape[1]
        H = D // 128
        Hl = H // P
        full = self.symm_buffer("full", (B, 3, P * Lp, Hl * 128), qkv.dtype, qkv.device)
        ops.qkv_norm_rope_scatter(qkv, D, w_q, w_k, eps, freqs, grid, lay.sp_rank * Lp, Lp, full.ptrs, Hl, P * Lp,
                                  lay.sp_rank * Lp)
        self._barrier(full)               # every rank's heads have landed before anyone attends over them
        f5 = full.tensor.view(B, 3, P * Lp, Hl, 128)
        esz = attn.tensor.element_size()
        head0 = lay.sp_rank * Hl * 128 * esz
        args = (f5[:, 0], f5[:, 1, :L], f5[:, 2, :L], [p + head0 for p in attn.ptrs], Lp, Lp * D, D, scale)
        if timed is None:
            ops.fmha_scatter(*args)
        else:
            timed("fmha", 4.0 * B * Hl * (P * Lp) * L * 128, ops.fmha_scatter, *args)
        self._barrier(full)               # every rank's rows have landed before the o projection reads them

    def _barrier(self, buf: "SymmBuffer") -> None:
        """Stream-ordered barrier over the SP group (symmetric-memory signal pads; a kernel, not a host wait)."""
        buf.handle.barrier(0)

    def _buf(self, name, shape, like):
        key = (name, tuple(shape), like.dtype)
        b = self._bufs.get(key)
        if b is None:
            b = torch.empty(tuple(shape), dtype=like.dtype, device=like.device)
            self._bufs[key] = b
        return b

    # -- Ulysses attention for ONE sample --------------------------------------------------------------------
    def attention(self, qkv: torch.Tensor, out: torch.Tensor, L: int, scale: float) -> None:
        """qkv: [Lp, 3, H, 128] local shard (q,k already normed + rotated); out: [Lp, H, 128]."""
        P = self.layout.sp_size
        Lp, _, H, hd = qkv.shape
        Hl = H // P
        inner = Hl * hd
        # [Lp][3*P][inner] -> [3*P][Lp][inner]: per (tensor, destination) send blocks
        send = self._buf("send", (3 * P, Lp, inner), qkv)
        self.swap01(qkv.view(Lp, 3 * P, inner), send)
        recv = self._buf("recv", (3, P * Lp, inner), qkv)
        for w in range(3):
            _all_to_all(recv[w], send[w * P:(w + 1) * P].view(P * Lp, inner), self.sp_group)
        full = recv.view(3, 1, P * Lp, Hl, hd)
        o_full = self._buf("o_full", (1, P * Lp, Hl, hd), qkv)
        # keys beyond the real sequence (SP padding) are masked by Lk = L; padded query rows are discarded below
        self.fmha(full[0], full[1][:, :L], full[2][:, :L], o_full, scale)
        o_recv = self._buf("o_recv", (P, Lp, inner), qkv)
        _all_to_all(o_recv.view(P * Lp, inner), o_full.view(P * Lp, inner), self.sp_group)
        # [P (head group)][Lp][inner] -> [Lp][P][inner] = [Lp, H, 128]
        self.swap01(o_recv, out.view(Lp, P, inner))

    def gather_tokens(self, local: torch.Tensor) -> torch.Tensor:
        """[Lp, C] -> [P*Lp, C] over the SP group (:1103-1104)."""
        P = self.layout.sp_size
        if P == 1:
            return local
        out = torch.empty((P * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        if dist.get_backend(self.sp_group) == "nccl":
            dist.all_gather_into_tensor(out, local.contiguous(), group=self.sp_group)
        else:
            parts = [torch.empty_like(local) for _ in range(P)]
            dist.all_gather(parts, local.contiguous(), group=self.sp_group)
            out.copy_(torch.cat(parts, 0))
        return out

    def fetch_from_last_cfg_rank(self, local: torch.Tensor) -> torch.Tensor:
        """Replace ``local`` by the same-shaped tensor of the LAST rank of this rank's CFG group (the cond branch):
        TeaCache residuals when cfg_skip turns CFG-branch parallelism off mid-run (NativeEngine._teacache_residual)."""
        src = self.layout.cfg_ranks(self.layout.sp_rank)[-1]
        buf = local.clone()
        dist.broadcast(buf, src=src, group=self.cfg_group)
        return buf

    def gather_cfg(self, local: torch.Tensor) -> torch.Tensor:
        """[b, ...] -> [cfg_size*b, ...] ordered by cfg_rank (uncond first, as the sampler built the batch)."""
        Cg = self.layout.cfg_size
        if Cg == 1:
            return local
        out = torch.empty((Cg * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        if dist.get_backend(self.cfg_group) == "nccl":
            dist.all_gather_into_tensor(out, local.contiguous(), group=self.cfg_group)
        else:
            parts = [torch.empty_like(local) for _ in range(Cg)]
            dist.all_gather(parts, local.contiguous(), group=self.cfg_group)
            out.copy_(torch.cat(parts, 0))
        return out


def setup(model, world: int, rank: int, cfg_size: Optional[int] = None) -> str:
    """Attach the partitioning to a native model (torch.distributed must be initialised). Returns e.g. 'cfg2xsp4'."""
    eng = model.engine() if hasattr(model, "engine") else model._flexam_engine
    layout = make_layout(world, rank, cfg_size, eng.H)
    eng.par = Parallel(layout, eng._swap01, eng._fmha)
    model.sp_world_size = layout.sp_size
    model.sp_world_rank = layout.sp_rank
    return layout.describe()


def attach(model) -> None:
    """``enable_multi_gpus_inference`` (:801-815): take the layout from the initialised default process group."""
    if not dist.is_initialized():
        raise RuntimeError("enable_multi_gpus_inference: torch.distributed is not initialised")
    setup(model, dist.get_world_size(), dist.get_rank())


def encode_many(vae, videos, world: Optional[int] = None, rank: Optional[int] = None):
    """The pipeline encodes 8 independent clips per generation (masked video, control video, additional controls,
    reference image; pipeline_wan2_2_fun_control_FlexAM.py:662-819): independent units, so across the ranks of one box
    they are sharded round-robin with no data-path collective except the final all-gather of the (small) latents.
    ``videos``: list of [1, 3, T, H, W] tensors, the same list on every rank (clips of one shape are gathered together).
    Returns the list of ``vae.encode(v).latent_dist.parameters`` in input order, identical on every rank."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
    mine = {i: vae.encode(v).latent_dist.parameters for i, v in enumerate(videos) if i % world == rank}
    if world == 1:
        return [mine[i] for i in range(len(videos))]
    out = [None] * len(videos)
    for i, v in enumerate(videos):          # one broadcast per clip from its owner (shapes are known on every rank)
        owner = i % world
        if owner == rank:
            buf = mine[i].contiguous()
        else:
            T = v.shape[2]
            buf = torch.empty((1, 2 * vae.cfg["z_dim"], 1 + (T - 1) // 4, v.shape[3] // 16, v.shape[4] // 16),
                              dtype=torch.bfloat16, device=v.device)
        dist.broadcast(buf, src=owner)
        out[i] = buf
    return out


# ----------------------------------------------------------------------------------------------------------
# VAE decode split across the ranks of one box
# ----------------------------------------------------------------------------------------------------------
class SlabExchange:
    """Band-of-rows decomposition of the VAE decode (``VaeDecoderEngine.slab``). The reference decodes the whole clip on
    every rank (pipeline_wan2_2_fun_control_FlexAM.py:955-958); here rank r decodes rows [r H / P, (r + 1) H / P) of the
    latent grid — and of every up-sampled stage — and the only coupling between bands is the one-row halo of each 3x3
    spatial convolution (RMS norm, SiLU, nearest up-sampling, the time convolution and the DupUp3D shortcut are per
    pixel) plus the single-head attention of the middle block, which every rank computes over the gathered frame. Every
    output pixel is produced by the same K-ordered accumulation as on one GPU, so the decoded clip is bit-identical.

    This base class moves halo rows with ``torch.distributed`` point-to-point calls (any backend: the world-size-2 gloo
    tests run it); ``SymmSlabExchange`` is the NVLink form (peer stores into symmetric buffers + a stream barrier)."""

    def __init__(self, world: int, rank: int, group=None):
        self.world, self.rank, self.group = world, rank, group
        self.exchanges = 0

    def alloc(self, name: str, rows: int, C: int, device) -> torch.Tensor:
        return torch.zeros((rows, C), dtype=torch.bfloat16, device=device)

    def sync(self) -> None:
        pass                                  # point-to-point receives land in private staging rows: nothing to order

    def halo(self, grid: torch.Tensor, frame0: int, T: int, Hp: int, Wp: int) -> None:
        g = grid[:(frame0 + T) * Hp * Wp].view(frame0 + T, Hp, Wp * grid.shape[1])[frame0:]
        up, dn = self.rank - 1, self.rank + 1
        ops_, recv = [], []
        if up >= 0:
            buf = torch.empty_like(g[:, 0])
            recv.append((buf, 0))
            ops_.append(dist.P2POp(dist.irecv, buf, self._global(up), self.group))
            ops_.append(dist.P2POp(dist.isend, g[:, 1].contiguous(), self._global(up), self.group))
        if dn < self.world:
            buf = torch.empty_like(g[:, 0])
            recv.append((buf, Hp - 1))
            ops_.append(dist.P2POp(dist.irecv, buf, self._global(dn), self.group))
            ops_.append(dist.P2POp(dist.isend, g[:, Hp - 2].contiguous(), self._global(dn), self.group))
        # one batch: under NCCL the sends and receives of a rank form one group (posted one by one, the receive-first
        # order of two neighbours would wait on each other)
        for r in dist.batch_isend_irecv(ops_):
            r.wait()
        for buf, row in recv:
            g[:, row].copy_(buf)
        self.exchanges += 1

    def _global(self, r: int) -> int:
        return r if self.group is None else dist.get_global_rank(self.group, r)

    def gather_rows(self, x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        dist.all_gather_into_tensor(out, x, group=self.group)
        return out

    def gather_video(self, v: torch.Tensor) -> torch.Tensor:
        """v: [3, T, band rows, W] -> [3, T, world * band rows, W]."""
        out = torch.empty((self.world * v.shape[0],) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
        dist.all_gather_into_tensor(out, v.contiguous(), group=self.group)           # rank-major along dim 0
        return out.view((self.world,) + tuple(v.shape)).permute(1, 2, 0, 3, 4).reshape(v.shape[0], v.shape[1], self.world * v.shape[2], v.shape[3])


class SymmSlabExchange(SlabExchange):
    """Grids in symmetric memory: one ``fx_vae_halo_push`` launch stores both boundary rows into the neighbours' grids over
    NVLink, then the buffer's own signal-pad barrier (a kernel on the stream, no host wait) orders it against their
    convolutions."""

    def __init__(self, world: int, rank: int, group=None):
        super().__init__(world, rank, group)
        self._h = {}
        self._names = {}
        self._sync_buf = None

    def _symm(self, shape, dtype, device):
        import torch.distributed._symmetric_memory as symm_mem
        t = symm_mem.empty(*shape, dtype=dtype, device=device)
        t.zero_()
        hdl = symm_mem.rendezvous(t, self.group if self.group is not None else dist.group.WORLD)
        hdl.barrier(0)                    # every rank has cleared its copy before any neighbour stores into it
        return t, hdl

    def alloc(self, name, rows, C, device):
        old = self._names.pop(name, None)
        if old is not None:
            self._h.pop(old, None)
        t, hdl = self._symm((rows, C), torch.bfloat16, device)
        self._h[t.data_ptr()] = (hdl, [int(p) for p in hdl.buffer_ptrs], t)
        self._names[name] = t.data_ptr()
        return t

    def sync(self):
        if self._sync_buf is None:
            dev = torch.device("cuda", torch.cuda.current_device())
            self._sync_buf = self._symm((64,), torch.float32, dev)
        self._sync_buf[1].barrier(0)

    def halo(self, grid, frame0, T, Hp, Wp):
        from . import ops
        hdl, ptrs, _ = self._h[grid.data_ptr()]
        up = ptrs[self.rank - 1] if self.rank > 0 else 0
        dn = ptrs[self.rank + 1] if self.rank + 1 < self.world else 0
        ops.vae_halo_push(grid, up, dn, frame0, T, Hp, Wp)
        hdl.barrier(0)
        self.exchanges += 1


def enable_vae_slabs(vae, world: Optional[int] = None, rank: Optional[int] = None, group=None):
    """Split ``vae.decode`` into bands of image rows over the ranks of ``group`` (default: all ranks). Every rank calls
    ``vae.decode(latents)`` with the same latents and gets the same, complete clip back. Latent heights that do not divide
    by the number of ranks fall back to the replicated decode."""
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    ex = None
    if world > 1:
        nccl = dist.get_backend(group) == "nccl" and os.environ.get("FLEXAM_VAE_HALO", "symm") != "p2p"
        ex = (SymmSlabExchange if nccl else SlabExchange)(world, rank, group)
    if hasattr(vae, "_slab"):
        vae._slab = ex                      # the mirror hands it to whichever engine it builds
    else:
        vae.slab = ex
    return ex
