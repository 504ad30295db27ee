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


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description="Image and attention map generation.")
    parser.add_argument("--save-dir", type=str, default="Data/Synthetic", help="Directory to save images (and heatmaps if enabled).")
    parser.add_argument("--pretrained-model-path", type=str, default="output/LINZ-Utah/sd1.4-token-finetune-stage-two/full_model_step_4500", help="Path or repo-id of the pretrained model to load.")
    parser.add_argument("--learnable-tokens-embedding-path", type=str, default="output/LINZ-Utah/sd1.4-token-finetune-stage-one/learned_embeds_steps_9000.bin", help="Path to the learned token embeddings.")
    parser.add_argument("--prompt", type=str, default="An aerial view image with {} cars in {} Utah", help="Prompt template for image generation.")
    parser.add_argument("--initialize_token", type=str, default=["cars", "Utah", "New Zealand"], nargs="+", help="The initialization for learnable tokens in the first stage")
    parser.add_argument("--word_token_heatmaps", type=str, default=None, nargs="+", help="word tokens to compute DAAM heatmaps.")
    parser.add_argument("--store_learnable_token_heatmaps", action="store_true", help="Whether to store DAAM heatmaps for learnable tokens.")
    parser.add_argument("--num-images", type=int, default=10000, help="Number of images to generate.")
    parser.add_argument("--image-size", type=int, default=112, help="Size of the generated images.")
    # additions
    parser.add_argument("--synthetic", action="store_true", help="Random-weight SD-1.5 attention stack instead of a diffusers pipeline.")
    parser.add_argument("--token-indices", type=int, nargs="+", default=None, help="Context-token rows of the words (needed when no tokenizer vocabulary is available).")
    parser.add_argument("--num-inference-steps", type=int, default=20)
    parser.add_argument("--batch-size", type=int, default=8, help="Images per batch in --synthetic mode.")
    return parser.parse_args(argv)


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


def run_pipeline(args):
    try:
        from diffusers import StableDiffusionPipeline
    except ImportError as e:  # pragma: no cover - not installable offline
        raise SystemExit("diffusers is not installed: real generation needs the reference's environment "
                         "(requirements.txt). Use --synthetic for the attention-stack workload.") from e
    from PIL import Image
    from . import trace as trace_mod
    os.makedirs(args.save_dir, exist_ok=True)
    pipeline = StableDiffusionPipeline.from_pretrained(args.pretrained_model_path).to("cuda")
    embeds_dict = torch.load(args.learnable_tokens_embedding_path)
    all_new_tokens = list(embeds_dict.keys())
    all_words = list(args.word_token_heatmaps) if args.word_token_heatmaps is not None else []
    new_tokens = []
    for t, n in zip(args.initialize_token, all_new_tokens):
        if t in args.prompt:
            if args.store_learnable_token_heatmaps:
                all_words.append(n)
            new_tokens.append(n)
    embeds = torch.stack([embeds_dict[token] for token in new_tokens]).to("cuda")
    pipeline.tokenizer.add_tokens(new_tokens)
    new_token_ids = pipeline.tokenizer.convert_tokens_to_ids(new_tokens)
    pipeline.text_encoder.resize_token_embeddings(len(pipeline.tokenizer))
    with torch.no_grad():
        pipeline.text_encoder.get_input_embeddings().weight.data[new_token_ids] = embeds
    prompt = args.prompt.format(*new_tokens)
    for seed in range(args.num_images):
        with trace_mod.trace(pipeline, prompt=prompt) as trc:
            generator = torch.Generator(device="cuda").manual_seed(seed)
            output_image = pipeline(prompt, num_inference_steps=args.num_inference_steps, generator=generator).images[0]
            output_image = output_image.resize((args.image_size, args.image_size))
            if np.max(np.asarray(output_image)) < 1e-5:  # NSFW content filter (data_generation.py:61-62)
                continue
            heat = trc.compute_global_heat_map()
        os.makedirs(os.path.join(args.save_dir, "images"), exist_ok=True)
        output_image.save(os.path.join(args.save_dir, "images", f"{seed}.png"))
        for word in all_words:
            hm = heat.compute_word_heat_map(word).heatmap
            save_word_heatmaps(args.save_dir, word, [seed], hm[None], args.image_size)


def main(argv=None):
    args = parse_args(argv)
    if not torch.cuda.is_available():
        raise SystemExit("agenda_b200.data_generation needs a CUDA device (there is no CPU path)")
    return run_synthetic(args) if args.synthetic else run_pipeline(args)


if __name__ == "__main__":
    main()
