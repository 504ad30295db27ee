"""`trace(pipeline)` — the daam-style context manager `data_generation.py:57-77` is written against:

    with trace(pipeline) as trc:
        pipeline(prompt, num_inference_steps=20, generator=g)
        heat = trc.compute_global_heat_map()
    heat.compute_word_heat_map(word).heatmap      # 2-D fp32 tensor [latent_hw, latent_hw]

so that swapping `import daam` for `from agenda_b200 import trace as daam` leaves the reference's generation loop
unchanged.  The third-party `daam` package is un-vendored and un-pinned (requirements.txt:4), cannot be installed
here, and no reference test pins its results: PARITY UNPINNED for the DAAM-specific choices below,
which are restated from daam's published algorithm:
  * only cross-attention (`attn2`) modules of the down and up blocks are hooked (mid block excluded): 15 layers;
  * batch-1 generation, the unconditional CFG half is dropped;
  * the result is truncated to len(tokenize(prompt)) + 2 rows when a tokenizer is available.
mode="hook" (default) aggregates with hook.py's semantics (mean over heads at native resolution, then bicubic+clamp
per map, mean over maps — SURVEY.md §8 a4), which IS pinned by the golden vectors; mode="daam" reproduces DAAM's own
aggregation as restated in oracle/hook_oracle.py:daam_global_heat_map: DAAM's layer selection, per-(layer, head) sums
over the denoising steps at native resolution, bicubic+clamp per (layer, head), mean over all of them.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from .processor import UNetCrossAttentionHooker


def word_token_indices(tokenizer, prompt: str, word: str) -> List[int]:
    """daam.utils.compute_token_merge_indices: positions (BOS offset +1) of the word's sub-token sequence in the
    lower-cased prompt (convention also at data_generation/dataset.py:93)."""
    if tokenizer is None:
        raise ValueError("no tokenizer: pass explicit token indices (compute_word_heat_map(word, token_idx=...))")
    split = tokenizer.tokenize if hasattr(tokenizer, "tokenize") else (lambda s: s.split())
    prompt_ids, word_ids = split(prompt.lower()), split(word.lower())
    hits = []
    for i in range(len(prompt_ids) - len(word_ids) + 1):
        if prompt_ids[i:i + len(word_ids)] == word_ids:
            hits += [i + 1 + k for k in range(len(word_ids))]
    if not hits:
        raise ValueError(f"Search word {word} not found in prompt!")
    return hits


class WordHeatMap:
    def __init__(self, heatmap: torch.Tensor, word: str):
        self.heatmap = heatmap  # [L, L] fp32 on the device, what data_generation.py:77-79 reads
        self.word = word

    @property
    def value(self):
        return self.heatmap


class GlobalHeatMap:
    def __init__(self, heat_maps: torch.Tensor, token_rows: Dict[int, int], tokenizer=None, prompt: str = ""):
        self.heat_maps = heat_maps  # [T, L, L]
        self._rows = token_rows     # context token index -> row of heat_maps
        self.tokenizer = tokenizer
        self.prompt = prompt

    def token_indices(self, word: str) -> List[int]:
        return word_token_indices(self.tokenizer, self.prompt, word)

    def compute_word_heat_map(self, word: str, token_idx: Optional[Sequence[int]] = None) -> WordHeatMap:
        idx = list(token_idx) if token_idx is not None else self.token_indices(word)
        rows = []
        for i in idx:
            if i not in self._rows:
                raise KeyError(f"token {i} was not traced (traced tokens: {sorted(self._rows)})")
            rows.append(self._rows[i])
        return WordHeatMap(self.heat_maps[rows].mean(0), word)


def _attention_modules(unet):
    """(qualified name, module) for every attention module that accepts a processor."""
    return [(n, m) for n, m in unet.named_modules() if hasattr(m, "set_processor") and hasattr(m, "to_q")]


class trace:
    def __init__(self, pipeline, tokens: Optional[Sequence[int]] = None, latent_hw: Optional[int] = None,
                 mode: str = "hook", precision: str = "bf16", prompt: str = ""):
        self.pipeline = pipeline
        self.unet = getattr(pipeline, "unet", pipeline)
        if latent_hw is None:  # daam: 64 for 512/1024-px pipelines, else 96
            cfg = getattr(self.unet, "config", None)
            sample = getattr(cfg, "sample_size", 64) if cfg is not None else 64
            latent_hw = int(sample) if isinstance(sample, int) and sample > 0 else 64
        self.mode = mode
        self.prompt = prompt
        self.tokens = None if tokens is None else list(tokens)
        self.hooker = UNetCrossAttentionHooker(is_train=False, latent_hw=latent_hw, tokens=self.tokens,
                                               precision=precision, aggregate="daam" if mode == "daam" else "hook")
        self._saved = []

    def __enter__(self):
        mods = _attention_modules(self.unet)
        if not mods:
            raise RuntimeError("trace(): no attention modules with set_processor() found on pipeline.unet")
        for name, m in mods:
            name = getattr(m, "block_name", None) or name
            is_cross = "attn2" in name or getattr(m, "is_cross", False) or getattr(m, "is_cross_attention", False)
            if self.mode == "daam" and (not is_cross or "mid" in name):
                continue  # DAAM hooks attn2 of the down/up blocks only; everything else keeps its processor
            self._saved.append((m, getattr(m, "processor", None)))
            m.set_processor(self.hooker)
        return self

    def __exit__(self, *exc):
        for m, proc in self._saved:
            m.set_processor(proc)
        self._saved.clear()
        return False

    def compute_global_heat_map(self, prompt: Optional[str] = None) -> GlobalHeatMap:
        try:
            heat = self.hooker.compute_global_heat_map()  # [B', T, L, L]
        except RuntimeError:
            raise RuntimeError('No heat maps found. Did you forget to call `with trace(...)` during generation?')
        heat = heat[0]  # batch-1 generation, as daam
        tokenizer = getattr(self.pipeline, "tokenizer", None)
        prompt = self.prompt if prompt is None else prompt
        n_rows = heat.shape[0]
        toks = self.tokens if self.tokens is not None else list(range(n_rows))
        if self.tokens is None and tokenizer is not None and prompt and hasattr(tokenizer, "tokenize"):
            keep = min(n_rows, len(tokenizer.tokenize(prompt)) + 2)
            heat, toks = heat[:keep], toks[:keep]
        return GlobalHeatMap(heat, {t: r for r, t in enumerate(toks)}, tokenizer, prompt)
