"""TEST INFRASTRUCTURE ONLY: compares the product's checkpoint-format tables (ltx/checkpoint_io.py) with the ones dumped from the
unmodified reference (oracle/gen_golden_formats.py -> tests/golden/ltx_format_tables.json)."""


UNUSED_REFERENCE_TRANSFORMER_KEYS = {"_diffusers_version", "_name_or_path", "attention_type", "double_self_attention", "dropout",
                                     "norm_num_groups", "num_embeds_ada_norm", "num_vector_embeds", "only_cross_attention",
                                     "project_to_2d_pos", "upcast_attention", "use_linear_projection"}


def check_format_tables(ref):
    from ltx_video_gpupoor_b200.ltx import checkpoint_io as C
    assert [list(kv) for kv in C.TRANSFORMER_KEYS_RENAME_DICT.items()] == ref["transformer_renames"]
    assert [list(kv) for kv in C.VAE_KEYS_RENAME_DICT.items()] == ref["vae_renames"]          # same entries in the same order
    assert C.DIFFUSERS_SCHEDULER_CONFIG == ref["diffusers_scheduler"]
    assert C.DIFFUSERS_TRANSFORMER_CONFIG == ref["diffusers_transformer"]
    assert C.DIFFUSERS_VAE_CONFIG == ref["diffusers_vae"]
    mapping = C.diffusers_and_ours_config_mapping()
    assert mapping[C.make_hashable_key(ref["diffusers_scheduler"])] == ref["ours_scheduler"]
    # transformer: every key both sides hold must agree; keys only the reference holds are diffusers leftovers its constructor accepts and
    # the path never reads; keys only the product holds must equal the reference constructor's defaults
    mine, theirs = mapping[C.make_hashable_key(ref["diffusers_transformer"])], ref["ours_transformer"]
    for k in set(mine) & set(theirs):
        assert mine[k] == theirs[k], k
    assert set(theirs) - set(mine) <= UNUSED_REFERENCE_TRANSFORMER_KEYS, set(theirs) - set(mine)
    for k in set(mine) - set(theirs):
        assert mine[k] == ref["transformer_init_defaults"][k], k
    assert mapping[C.make_hashable_key(ref["diffusers_vae"])] == ref["ours_vae"]
    assert repr(C.make_hashable_key(ref["hashable_probe"]["arg"])) == ref["hashable_probe"]["repr"]
