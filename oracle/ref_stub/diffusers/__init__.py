"""TEST INFRASTRUCTURE ONLY — minimal stand-in for the `diffusers` package.

`/root/reference/data_generation/hook.py` imports `diffusers.StableDiffusionPipeline` (unused by the
hooker) and `diffusers.models.attention_processor.Attention` (type hint; the hooker only calls methods on
the module it is handed).  diffusers is not installed in this image and there is no network, so the golden
generator (`oracle/gen_golden.py`) puts this directory on `sys.path` to execute hook.py UNMODIFIED.
Nothing under `agenda_b200/` may import this.
"""


class StableDiffusionPipeline:  # hook.py:5 imports the name only
    pass
