"""The hot path end to end, as one callable: (UNet attention hidden states, prompt context) -> heat maps ->
u8 stacks + boxes.  This is the public API `bench.py` measures and the generation driver
(`agenda_b200/data_generation.py`) uses.

One `HeatmapPipeline.run_*` call == what the reference does for a batch of images between
`with daam.trace(pipeline)` and the PNG writes (data_generation.py:57-86) plus `postprocess_heatmap.py:44-46`,
with the attention processor of hook.py:83-122 doing the capture — minus the non-attention UNet layers, which
are out of scope (SURVEY.md §8 f N4): the 32 attention calls of every denoising step are fed synthetic hidden
states of the right shapes.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import ops
from .processor import UNetCrossAttentionHooker
from .mixed import compensate_cross_projections
from .sd_attention import AttentionStack, BlockSpec, sd15_blocks, sd21_blocks


class HeatmapPipeline:
    def __init__(self, blocks: Optional[List[BlockSpec]] = None, context_dim: int = 768,
                 tokens: Sequence[int] = (5, 6, 7), num_steps: int = 50, latent_hw: int = 64, image_size: int = 112,
                 dtype: torch.dtype = torch.bfloat16, device="cuda", thr: float = 0.5, max_boxes: int = 64,
                 seed: int = 0, precision: str = "bf16", use_cuda_graph: bool = True, cross_logits: str = "fp32",
                 compensate_weights: bool = True):
        self.device = torch.device(device)
        self.blocks = blocks if blocks is not None else sd15_blocks(latent_hw)
        # the "checkpoint" is fp32 (AttentionStack(blocks, context_dim, seed) reproduces it for a reference run); a 16-bit
        # pipeline keeps the fp32 value of the projections that feed the heat-map logits as weight + weight_lo
        self._weight_seed = seed
        stack = AttentionStack(self.blocks, context_dim, seed)
        if dtype != torch.float32 and compensate_weights and cross_logits == "fp32":
            compensate_cross_projections(stack)
        self.stack = stack.to(device=self.device, dtype=dtype)
        self.tokens = list(tokens)
        if len(self.tokens) < 3:
            raise ValueError("need at least (object, fg, bg) token indices")
        self.num_steps = num_steps
        self.latent_hw = latent_hw
        self.image_size = image_size
        self.dtype = dtype
        self.thr = thr
        self.max_boxes = max_boxes
        self.proc = UNetCrossAttentionHooker(is_train=False, latent_hw=latent_hw, tokens=self.tokens,
                                             precision=precision, cross_logits=cross_logits)
        self.stack.set_attn_processor(self.proc)
        self.use_cuda_graph = use_cuda_graph
        self._graph = None
        self._graph_key = None

    def reference_stack(self) -> AttentionStack:
        """The fp32 "checkpoint" this pipeline was built from, on the CPU (deterministic in (blocks, context_dim, seed)):
        what a reference run, or the oracle in the tests, uses."""
        return AttentionStack(self.blocks, self.stack.context_dim, self._weight_seed)

    # ------------------------------------------------------------------------------------------------------
    def make_inputs(self, n_images: int, seed: int = 0, pinned_host: bool = False, seeds: Optional[Sequence[int]] = None,
                    on_device: bool = False):
        """Synthetic inputs for `n_images` (UNet batch = 2*n_images with classifier-free guidance).  With `seeds` (one
        per image) every image's inputs depend on its own seed only, whatever batch or rank it lands in; on_device=True
        then draws them on the GPU (returns device tensors)."""
        if seeds is not None:
            if len(seeds) != n_images:
                raise ValueError("need one seed per image")
            if on_device:
                return self.stack.make_inputs_for_seeds(seeds, self.device, self.dtype, on_device=True)
            hs, ctx = self.stack.make_inputs_for_seeds(seeds, "cpu", self.dtype)
        else:
            hs, ctx = self.stack.make_inputs(2 * n_images, "cpu", self.dtype, seed)
        if pinned_host:
            return {k: v.pin_memory() for k, v in hs.items()}, ctx.pin_memory()
        return {k: v.to(self.device) for k, v in hs.items()}, ctx.to(self.device)

    @torch.no_grad()
    def run_device(self, hs: Dict, ctx: torch.Tensor) -> Dict[str, torch.Tensor]:
        """All tensors on the device.  Returns heat [n,T,L,L] fp32, planes u8 [n,3,S,S], stack u8 [n,S,S,3], inv u8
        [n,S,S], counts int32 [n], boxes int32 [n,max_boxes,5] (x,y,w,h,area on the object-token map)."""
        proc = self.proc
        proc.clear(keep_context_kv=True)  # the captured graph reads the cached prompt K/V buffers
        # every run is a new batch of images: the prompt K/V are projected once per batch here (not once per denoising
        # step as in the reference, hook.py:101-102) — also picks up an embedding overwritten in place since the last run
        proc.refresh_context_kv(force=True)
        key = (tuple((k, v.data_ptr(), tuple(v.shape), v.dtype) for k, v in sorted(hs.items())), ctx.data_ptr(),
               tuple(ctx.shape), ctx.dtype)
        if self.use_cuda_graph:
            if self._graph is None or self._graph_key != key:
                # warm the allocator/cuBLAS handles outside capture, then capture one denoising step's 32 calls
                proc.clear()             # new input buffers: rebuild the prompt K/V cache outside the capture
                self.stack(hs, ctx)
                proc.clear(keep_context_kv=True)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                from . import _lib
                l0 = _lib.launches
                with torch.cuda.graph(g):
                    self.stack(hs, ctx)
                self._graph_launches = _lib.launches - l0  # C-ABI kernel nodes captured per denoising step
                self._graph, self._graph_key = g, key
                self._maps_per_step = proc.num_maps
                proc.clear(keep_context_kv=True)
            for _ in range(self.num_steps):
                self._graph.replay()
            proc._count = self._maps_per_step * self.num_steps
        else:
            for _ in range(self.num_steps):
                self.stack(hs, ctx)
        heat = proc.compute_global_heat_map()                       # hook.py:59-81
        planes, stack, inv = ops.heat_postprocess_stack(heat[:, :3].contiguous(), self.image_size)
        _, counts, boxes = ops.ccl_bbox(heat[:, 0].contiguous(), self.thr, self.max_boxes, want_labels=False)
        return {"heat": heat, "planes": planes, "stack": stack, "inv": inv, "counts": counts, "boxes": boxes}

    @torch.no_grad()
    def run_host(self, hs_host: Dict, ctx_host: torch.Tensor, staging: Optional[Dict] = None) -> Dict:
        """Host buffers in (pinned), host buffers out: H2D of the step's inputs, the hot path, D2H of the results.
        `staging` (from make_staging) provides persistent device/pinned buffers so the CUDA graph stays valid."""
        if staging is None:
            staging = self.make_staging(hs_host, ctx_host)
        for k, v in hs_host.items():
            staging["hs"][k].copy_(v, non_blocking=True)
        staging["ctx"].copy_(ctx_host, non_blocking=True)
        out = self.run_device(staging["hs"], staging["ctx"])
        self.last_device_out = out
        host = staging["out"]
        for name in ("stack", "inv", "planes", "counts", "boxes", "heat"):
            if name not in host:
                host[name] = torch.empty(out[name].shape, dtype=out[name].dtype).pin_memory()
            host[name].copy_(out[name], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host

    @torch.no_grad()
    def run_host_batches(self, batches, staging: Optional[Dict] = None):
        """A stream of host batches `(hs_host, ctx_host)` (pinned) in, one dict of pinned host records out per batch
        (a generator; the record buffers are re-used, consume them before asking for the next).  Same work per batch as
        `run_host`, but the H2D copy of batch i+1 runs on a copy stream into a second set of device buffers while batch
        i computes; a device-to-device copy (tens of microseconds) then moves it into the buffers the captured CUDA graph
        reads.  Only the first batch's upload is exposed — on a box with a slow host link `run_host` spends 15 % of
        its time there."""
        it = iter(batches)
        cur = next(it, None)
        if cur is None:
            return
        if staging is None:
            staging = self.make_staging(*cur)
        if "hs_next" not in staging:
            staging["hs_next"] = {k: torch.empty_like(v) for k, v in staging["hs"].items()}
            staging["ctx_next"] = torch.empty_like(staging["ctx"])
            staging["copy_stream"] = torch.cuda.Stream(device=self.device)
        main, side = torch.cuda.current_stream(self.device), staging["copy_stream"]
        uploaded, consumed = torch.cuda.Event(), torch.cuda.Event()

        def upload(batch):
            hs_host, ctx_host = batch
            with torch.cuda.stream(side):
                for k, v in hs_host.items():
                    staging["hs_next"][k].copy_(v, non_blocking=True)
                staging["ctx_next"].copy_(ctx_host, non_blocking=True)
                uploaded.record(side)

        side.wait_stream(main)
        upload(cur)
        while cur is not None:
            main.wait_event(uploaded)
            for k, v in staging["hs_next"].items():
                staging["hs"][k].copy_(v, non_blocking=True)
            staging["ctx"].copy_(staging["ctx_next"], non_blocking=True)
            consumed.record(main)
            nxt = next(it, None)
            if nxt is not None:
                side.wait_event(consumed)
                upload(nxt)
            out = self.run_device(staging["hs"], staging["ctx"])
            self.last_device_out = out
            host = staging["out"]
            for name in ("stack", "inv", "planes", "counts", "boxes", "heat"):
                if name not in host:
                    host[name] = torch.empty(out[name].shape, dtype=out[name].dtype).pin_memory()
                host[name].copy_(out[name], non_blocking=True)
            main.synchronize()
            yield host
            cur = nxt

    @torch.no_grad()
    def run_seeds(self, seeds: Sequence[int], batch_size: int = 8, staging: Optional[Dict] = None) -> Dict[str, torch.Tensor]:
        """The sharded-generation loop of one rank (data_generation.py:56-59 over this rank's seeds): batches of
        `batch_size` images, inputs drawn on the device from each image's own seed, persistent staging buffers (one CUDA
        graph for the whole run), a short last batch padded with repeats and trimmed.  Returns per-image records in
        `seeds` order: heat [n,T,L,L] fp32, stack u8 [n,S,S,3], counts int32 [n], boxes int32 [n,max_boxes,5]."""
        keep = ("heat", "stack", "counts", "boxes")
        parts = {k: [] for k in keep}
        seeds = list(seeds)
        for i in range(0, len(seeds), batch_size):
            chunk = seeds[i:i + batch_size]
            n = len(chunk)
            hs, ctx = self.make_inputs(batch_size, seeds=chunk + [chunk[-1]] * (batch_size - n), on_device=True)
            if staging is None:
                staging = {"hs": {k: torch.empty_like(v) for k, v in hs.items()}, "ctx": torch.empty_like(ctx)}
            for k, v in hs.items():
                staging["hs"][k].copy_(v)
            staging["ctx"].copy_(ctx)
            out = self.run_device(staging["hs"], staging["ctx"])
            for k in keep:
                parts[k].append(out[k][:n].clone())
        self._seed_staging = staging
        if not seeds:
            L, S, T = self.latent_hw, self.image_size, len(self.tokens)
            return {"heat": torch.empty((0, T, L, L), device=self.device), "stack": torch.empty((0, S, S, 3), dtype=torch.uint8, device=self.device),
                    "counts": torch.empty((0,), dtype=torch.int32, device=self.device),
                    "boxes": torch.empty((0, self.max_boxes, 5), dtype=torch.int32, device=self.device)}
        return {k: torch.cat(v, 0) for k, v in parts.items()}

    def make_staging(self, hs_host: Dict, ctx_host: torch.Tensor) -> Dict:
        return {"hs": {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in hs_host.items()},
                "ctx": torch.empty(ctx_host.shape, dtype=ctx_host.dtype, device=self.device), "out": {}}

    @staticmethod
    def h2d_bytes(hs_host: Dict, ctx_host: torch.Tensor) -> int:
        return sum(v.numel() * v.element_size() for v in hs_host.values()) + ctx_host.numel() * ctx_host.element_size()

    @staticmethod
    def d2h_bytes(host_out: Dict) -> int:
        return sum(v.numel() * v.element_size() for v in host_out.values())


def sd15_pipeline(**kw) -> HeatmapPipeline:
    return HeatmapPipeline(sd15_blocks(kw.pop("latent_hw", 64)), 768, latent_hw=64, **kw)


def sd21_pipeline(**kw) -> HeatmapPipeline:
    return HeatmapPipeline(sd21_blocks(96), 1024, latent_hw=96, **kw)
