"""Checkpoint-format tables (data, not arithmetic): the reference's diffusers <-> native config mapping and key-rename tables
(ltx_video/utils/diffusers_config_mapping.py:1-174) and its scheduler / transformer / VAE configs, dumped from the UNMODIFIED
reference module into tests/golden/ltx_format_tables.json so that the product's copies in ltx/checkpoint_io.py are checked
entry by entry AND in order (the renames are applied as successive str.replace calls, so order is part of the format).
TEST INFRASTRUCTURE ONLY.  Build container only (needs /root/reference):  python oracle/gen_golden_formats.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "refshim"))
import load_reference  # noqa: E402

load_reference.install()


def main():
    import ltx_video.utils.diffusers_config_mapping as M
    import inspect
    from ltx_video.models.transformers.transformer3d import Transformer3DModel
    defaults = {k: v.default for k, v in inspect.signature(Transformer3DModel.__init__).parameters.items()
                if v.default is not inspect.Parameter.empty and isinstance(v.default, (int, float, str, bool, type(None), list, tuple))}
    out = dict(
        transformer_init_defaults=defaults,
        transformer_renames=list(M.TRANSFORMER_KEYS_RENAME_DICT.items()), vae_renames=list(M.VAE_KEYS_RENAME_DICT.items()),
        diffusers_scheduler=M.DIFFUSERS_SCHEDULER_CONFIG, diffusers_transformer=M.DIFFUSERS_TRANSFORMER_CONFIG,
        diffusers_vae=M.DIFFUSERS_VAE_CONFIG, ours_scheduler=M.OURS_SCHEDULER_CONFIG, ours_transformer=M.OURS_TRANSFORMER_CONFIG,
        ours_vae=M.OURS_VAE_CONFIG,
        # make_hashable_key on a nested probe (lists -> tuples, dicts -> sorted item tuples), stored as its repr
        hashable_probe=dict(arg={"b": [1, 2], "a": {"y": [3], "x": 1}, "c": "s"},
                            repr=repr(M.make_hashable_key({"b": [1, 2], "a": {"y": [3], "x": 1}, "c": "s"}))))
    from oracle.format_tables import check_format_tables
    check_format_tables(json.loads(json.dumps(out)))
    path = os.path.join(ROOT, "tests", "golden", "ltx_format_tables.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=False)
    print("written", path)
    trim_sequence_grid()


def trim_sequence_grid():
    """LTXVideoPipeline.trim_conditioning_sequence (pipeline_ltx_video.py:1689-1707): the reference method on a grid of
    (start frame, sequence length, video length) -> tests/golden/ltx_trim_sequence.json; the product method must agree on every row."""
    from types import SimpleNamespace
    import ltx_video.pipelines.pipeline_ltx_video as R
    from ltx_video_gpupoor_b200.ltx.pipeline_ltx_video import LTXVideoPipeline
    ours = LTXVideoPipeline.__new__(LTXVideoPipeline)
    ours.video_scale_factor = 8
    ref = SimpleNamespace(video_scale_factor=8)
    rows = []
    for start in (0, 8, 16, 24, 40, 96):
        for n in (1, 9, 17, 25, 33, 49, 57, 97, 121, 130):
            for target in (9, 17, 33, 97, 121, 257):
                if target - start < 1:
                    continue
                r = R.LTXVideoPipeline.trim_conditioning_sequence(ref, start, n, target)
                assert ours.trim_conditioning_sequence(start, n, target) == r
                rows.append([start, n, target, r])
    with open(os.path.join(ROOT, "tests", "golden", "ltx_trim_sequence.json"), "w") as f:
        json.dump(rows, f)
    print("written tests/golden/ltx_trim_sequence.json:", len(rows), "rows identical")


if __name__ == "__main__":
    main()
