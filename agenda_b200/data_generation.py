"""CLI mirror of the reference's `data_generation/data_generation.py` (same flags, same output tree:
`images/{seed}.png`, `daam_{word}_heatmaps/{seed}.png`), with the heat-map path on the B200 kernels:

  * `daam.trace(pipeline)` (data_generation.py:57)            -> `agenda_b200.trace.trace(pipeline, tokens=...)`
  * `.compute_word_heat_map(word).heatmap` (:74-77)           -> same call on the shim
  * min-max -> u8 -> PIL resize (:78-85, numpy/PIL on the CPU) -> `agenda_heat_to_u8_image` on the GPU

Two modes:
  * real generation (needs `diffusers` + a checkpoint; neither exists in this offline image, so this branch is
    exercised only by maintainers with the reference's environment);
  * `--synthetic`: the SD-1.5 attention stack with random weights and synthetic hidden states (the workload
    bench.py measures), sharded by seed over the visible GPUs when launched with torchrun.  No RGB image is
    produced in this mode (there is no VAE); heat-map PNGs are.
"""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch


# The reference's command line (data_generation.py:11-23): flag -> (type, default, extra argparse keywords).  Names
# and defaults are the drop-in surface; the help texts below are ours.
_REFERENCE_FLAGS = (
    ("--save-dir", str, "Data/Synthetic", {}, "output root: images/ and daam_<word>_heatmaps/ are created below it"),
    ("--pretrained-model-path", str, "output/LINZ-Utah/sd1.4-token-finetune-stage-two/full_model_step_4500", {},
     "Stable Diffusion checkpoint (directory or hub id)"),
    ("--learnable-tokens-embedding-path", str,
     "output/LINZ-Utah/sd1.4-token-finetune-stage-one/learned_embeds_steps_9000.bin", {},
     "file with the embeddings of the learned tokens"),
    ("--prompt", str, "An aerial view image with {} cars in {} Utah", {}, "prompt template; {} = learned tokens"),
    ("--initialize_token", str, ["cars", "Utah", "New Zealand"], {"nargs": "+"},
     "words the learned tokens were initialised from (stage one)"),
    ("--word_token_heatmaps", str, None, {"nargs": "+"}, "ordinary words whose heat maps are wanted"),
    ("--num-images", int, 10000, {}, "how many seeds to run"),
    ("--image-size", int, 112, {}, "side of the saved images and heat maps"),
)


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description="Heat-map-labelled image generation on the B200 kernels.")
    for flag, typ, default, extra, text in _REFERENCE_FLAGS:
        parser.add_argument(flag, type=typ, default=default, help=text, **extra)
    parser.add_argument("--store_learnable_token_heatmaps", action="store_true",
                        help="also save the heat maps of the learned tokens")
    # not in the reference
    parser.add_argument("--synthetic", action="store_true",
                        help="random-weight SD-1.5 attention stack instead of a diffusers pipeline")
    parser.add_argument("--token-indices", type=int, nargs="+", default=None,
                        help="context rows of the words (when no tokenizer vocabulary is available)")
    parser.add_argument("--num-inference-steps", type=int, default=20)
    parser.add_argument("--batch-size", type=int, default=8, help="images per batch in --synthetic mode")
    return parser


def parse_args(argv=None):
    return build_parser().parse_args(argv)


def save_word_heatmaps(save_dir, word, seeds, heat, image_size):
    """data_generation.py:71-86 for a batch: heat fp32 [n,L,L] (device) -> daam_{word}_heatmaps/{seed}.png."""
    from PIL import Image
    from . import ops
    out_dir = os.path.join(save_dir, "daam_" + word + "_heatmaps")
    os.makedirs(out_dir, exist_ok=True)
    u8 = ops.heat_to_u8_image(heat.contiguous(), image_size).cpu().numpy()
    for k, seed in enumerate(seeds):
        Image.fromarray(u8[k]).save(os.path.join(out_dir, f"{seed}.png"))


def run_synthetic(args):
    import torch.distributed as dist
    from .pipeline import sd15_pipeline
    from .sharding import shard_seeds
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    words = list(args.word_token_heatmaps or ["cars", "fg", "bg"])
    toks = args.token_indices or list(range(5, 5 + len(words)))
    if len(toks) != len(words):
        raise SystemExit("--token-indices must give one context row per word")
    while len(toks) < 3:
        toks = toks + [toks[-1] + 1]
    pipe = sd15_pipeline(tokens=toks, num_steps=args.num_inference_steps, device=f"cuda:{local}")
    seeds = shard_seeds(args.num_images, rank, world)
    os.makedirs(args.save_dir, exist_ok=True)
    done = 0
    for i in range(0, len(seeds), args.batch_size):
        chunk = seeds[i:i + args.batch_size]
        n = len(chunk)
        hs, ctx = pipe.make_inputs(args.batch_size, seed=chunk[0])
        out = pipe.run_device(hs, ctx)
        for w, word in enumerate(words):
            save_word_heatmaps(args.save_dir, word, chunk, out["heat"][:n, w], args.image_size)
        done += n
    if world > 1:
        dist.barrier()
    return done


def _install_learned_tokens(pipeline, args):
    """data_generation.py:40-54: register the learned tokens that the prompt template uses with the tokenizer and copy
    their embeddings into the text encoder.  Returns (token strings used in the prompt, words to draw heat maps for)."""
    learned = torch.load(args.learnable_tokens_embedding_path)
    used, words = [], list(args.word_token_heatmaps or [])
    for init_word, token in zip(args.initialize_token, learned.keys()):
        if init_word not in args.prompt:
            continue
        used.append(token)
        if args.store_learnable_token_heatmaps:
            words.append(token)
    tok, enc = pipeline.tokenizer, pipeline.text_encoder
    tok.add_tokens(used)
    enc.resize_token_embeddings(len(tok))
    rows = tok.convert_tokens_to_ids(used)
    with torch.no_grad():
        table = enc.get_input_embeddings().weight
        table.data[rows] = torch.stack([learned[t] for t in used]).to(table.device, table.dtype)
    return used, words


def run_pipeline(args):
    try:
        from diffusers import StableDiffusionPipeline
    except ImportError as e:  # pragma: no cover - not installable offline
        raise SystemExit("diffusers is not installed: real generation needs the reference's environment "
                         "(requirements.txt). Use --synthetic for the attention-stack workload.") from e
    from . import trace as trace_mod
    pipeline = StableDiffusionPipeline.from_pretrained(args.pretrained_model_path).to("cuda")
    used, words = _install_learned_tokens(pipeline, args)
    prompt = args.prompt.format(*used)
    image_dir = os.path.join(args.save_dir, "images")
    os.makedirs(image_dir, exist_ok=True)
    for seed in range(args.num_images):
        rng = torch.Generator(device="cuda").manual_seed(seed)
        with trace_mod.trace(pipeline, prompt=prompt) as tracer:          # daam.trace(pipeline), :57
            image = pipeline(prompt, num_inference_steps=args.num_inference_steps, generator=rng).images[0]
            image = image.resize((args.image_size, args.image_size))
            if np.asarray(image).max() < 1e-5:                            # all-black = filtered image, :61-62
                continue
            heat = tracer.compute_global_heat_map()
        image.save(os.path.join(image_dir, f"{seed}.png"))
        for word in words:
            plane = heat.compute_word_heat_map(word).heatmap              # :74-77
            save_word_heatmaps(args.save_dir, word, [seed], plane[None], args.image_size)
    return args.num_images


def main(argv=None):
    args = parse_args(argv)
    if not torch.cuda.is_available():
        raise SystemExit("agenda_b200.data_generation needs a CUDA device (there is no CPU path)")
    return run_synthetic(args) if args.synthetic else run_pipeline(args)


if __name__ == "__main__":
    main()
