"""CPU (gloo, world_size 2): the Ulysses pack / all-to-all / unpack plumbing of the sequence-parallel path
reproduces single-rank attention exactly, and the oracle's virtual-rank emulation agrees with it."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import wan_oracle as W
from oracle.ltx_oracle import attention_core


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ltx_video_gpupoor_b200.wan.distributed.ulysses import ulysses_self_attention
        B, N, H, d = 2, 24, 4, 8
        g = torch.Generator().manual_seed(0)
        qkv = torch.randn(B, N, 3 * H * d, generator=g)
        n_loc = N // world
        local = qkv[:, rank * n_loc:(rank + 1) * n_loc].reshape(B * n_loc, 3 * H * d).contiguous()

        def attn(q, k, v, out):
            out.copy_(attention_core(q.contiguous(), k.contiguous(), v.contiguous()))

        o = ulysses_self_attention(local, B, n_loc, H, d, dist.group.WORLD, attn)
        q, k, v = [qkv[:, :, i * H * d:(i + 1) * H * d].reshape(B, N, H, d) for i in range(3)]
        ref = attention_core(q, k, v).reshape(B, N, H * d)[:, rank * n_loc:(rank + 1) * n_loc].reshape(B * n_loc, H * d)
        ret[rank] = float((o - ref).abs().max())
        virt = W.ulysses_attention_virtual(q, k, v, world).reshape(B, N, H * d)
        ret[rank + world] = float((virt - attention_core(q, k, v).reshape(B, N, H * d)).abs().max())
    finally:
        dist.destroy_process_group()


def test_ulysses_exchange_two_ranks_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == 2 * world
    for k, v in ret.items():
        assert v < 1e-5, (k, v)


def _cfgp_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ltx_video_gpupoor_b200.wan.distributed.cfg_parallel import CfgParallel
        cp = CfgParallel()
        assert cp.branch == rank // (world // 2) and dist.get_world_size(cp.sp_group) == world // 2
        assert cp.select("cond", "uncond") == ("cond" if rank < world // 2 else "uncond")
        mine = torch.full((3, 5), float(10 * cp.branch + 1))          # cond half predicts 1, uncond half 11
        c, u = cp.exchange(mine)
        ret[rank] = (float(c.mean()), float(u.mean()))
    finally:
        dist.destroy_process_group()


def test_cfg_parallel_groups_and_exchange_gloo():
    """CfgParallel (distributed/cfg_parallel.py): two branch groups, pairwise exchange ordered (cond, uncond) on every rank."""
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_cfgp_worker, args=(world, 31500 + (os.getpid() % 2000), ret), nprocs=world, join=True)
    assert dict(ret) == {0: (1.0, 11.0), 1: (1.0, 11.0)}


def test_pack_unpack_are_pure_permutations():
    from ltx_video_gpupoor_b200.wan.distributed import ulysses as U
    B, n_loc, P, H, d = 2, 3, 2, 4, 8
    qkv = torch.arange(B * n_loc * 3 * H * d, dtype=torch.float32).reshape(B * n_loc, 3 * H * d)
    send = U.pack_qkv(qkv, B, n_loc, P, H, d)
    assert send.shape == (P, n_loc, B, 3, H // P, d)
    assert torch.equal(send.flatten().sort().values, qkv.flatten().sort().values)
    # peer p receives exactly the columns of its head group
    q = qkv[:, : H * d].reshape(B, n_loc, H, d)
    assert torch.equal(send[1, :, :, 0], q[:, :, 2:4].permute(1, 0, 2, 3))


def test_unipc_host_scalars_match_oracle():
    from ltx_video_gpupoor_b200.wan.fm_solvers_unipc import FlowUniPCMultistepScheduler
    for steps, shift in ((4, 5.0), (50, 5.0), (9, 3.0)):
        s = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
        s.set_timesteps(steps, device="cpu", shift=shift)
        o = W.UniPC(); o.set_timesteps(steps, shift)
        assert torch.equal(s.timesteps, o.timesteps) and torch.equal(s.sigmas, o.sigmas)


def test_wan_rope_tables_bit_exact(golden_dir):
    from ltx_video_gpupoor_b200.wan.posemb_layers import get_rotary_pos_embed
    g = torch.load(os.path.join(golden_dir, "wan_t2v.pt"), weights_only=False)
    cos, sin = get_rotary_pos_embed((3, 8, 12))
    c2, s2 = W.rope_tables((3, 8, 12))
    assert torch.equal(cos, c2) and torch.equal(sin, s2)
    assert torch.equal(cos[17], g["cos_row"]) and torch.equal(sin[17], g["sin_row"])


def test_wan_oracle_matches_reference_fixture(golden_dir):
    g = torch.load(os.path.join(golden_dir, "wan_t2v.pt"), weights_only=False)
    cfg = g["cfg"]
    sd = {k: v.double() for k, v in W.make_wan_state_dict(cfg, seed=0).items()}
    cos, sin = W.rope_tables(g["lat"].shape[1:])
    y = W.wan_forward(sd, cfg, [g["lat"].double(), g["lat"].double()], g["t"], [g["ctx"].double(), g["ctx0"].double()], cos, sin)
    for a, b in zip(y, g["fwd"]):
        assert W.rel_l2(a, b) < 2e-5
    steps = []
    W.t2v_denoise(sd, cfg, g["lat"].double(), g["ctx"].double(), g["ctx0"].double(), steps=4, shift=5.0, guide_scale=5.0, per_step=steps)
    for a, b in zip(steps, g["loop"]):
        assert W.rel_l2(a, b) < 5e-5


def test_wan_rope_tables_riflex_bit_exact(golden_dir):
    """get_rotary_pos_embed(enable_RIFLEx=...) against rows recorded from the unmodified reference (oracle/gen_golden_rope.py)."""
    from ltx_video_gpupoor_b200.wan.posemb_layers import get_rotary_pos_embed
    g = torch.load(os.path.join(golden_dir, "wan_rope_riflex.pt"), weights_only=False)
    assert len(g) == 4
    for (size, rf), (c_ref, s_ref) in g.items():
        cos, sin = get_rotary_pos_embed(size, enable_RIFLEx=rf)
        assert torch.equal(cos[::7], c_ref) and torch.equal(sin[::7], s_ref), (size, rf)
    a, b = get_rotary_pos_embed((33, 8, 12), False), get_rotary_pos_embed((33, 8, 12), True)
    assert not torch.equal(a[0][:, :44], b[0][:, :44]) and torch.equal(a[0][:, 44:], b[0][:, 44:])      # only the time axis changes


def test_wan_1_3b_full_depth_oracle_matches_reference_fixture(golden_dir):
    """BASELINE.json configs[3]'s network at full width and depth (dim 1536, 12 heads, ffn 8960, 30 layers, 1.42 B seeded weights): the oracle
    in fp32 against the joint forward and the 3-step CFG loop recorded from the unmodified reference in fp64 (oracle/gen_golden_wan_full.py)."""
    g = torch.load(os.path.join(golden_dir, "wan_1_3b_full.pt"), weights_only=False)
    cfg = g["cfg"]
    assert (cfg["dim"], cfg["num_heads"], cfg["ffn_dim"], cfg["num_layers"]) == (1536, 12, 8960, 30)
    sd = W.make_wan_state_dict(cfg, seed=g["seed_weights"])
    cos, sin = W.rope_tables(g["lat"].shape[1:])
    with torch.no_grad():
        y = W.wan_forward(sd, cfg, [g["lat"], g["lat"]], g["t"], [g["ctx"], g["ctx0"]], cos, sin)
        for a, b in zip(y, g["fwd"]):
            assert W.rel_l2(a, b) < 2e-4              # fp32 oracle vs fp64 reference through 30 layers
        steps = []
        W.t2v_denoise(sd, cfg, g["lat"], g["ctx"], g["ctx0"], steps=g["steps"], shift=g["shift"], guide_scale=g["guide"], per_step=steps)
    for a, b in zip(steps, g["loop"]):
        assert W.rel_l2(a, b) < 2e-4


def test_oracle_loop_matches_reference_generate_method(golden_dir):
    """The oracle's denoise loop vs the latents the reference's OWN WanT2V.generate method returned (oracle/gen_golden_wan_generate.py: the
    unmodified method called on a stand-in self): noise from `seed`, UniPC / dpm++, CFG-Zero* around cfg_zero_step, joint and two-call
    passes, skip-layer guidance on a window of steps, guide_scale == 1."""
    g = torch.load(os.path.join(golden_dir, "wan_generate.pt"), weights_only=False)
    cfg = g["cfg"]
    sd = {k: v.double() for k, v in W.make_wan_state_dict(cfg, seed=0).items()}
    assert len(g["cases"]) == 5
    for name, c in g["cases"].items():
        kw = c["kw"]
        noise = torch.randn(16, 3, 8, 12, dtype=torch.float32, generator=torch.Generator().manual_seed(c["seed"]))
        with torch.no_grad():
            lat = W.t2v_denoise(sd, cfg, noise.double(), g["ctx"].double(), g["ctx0"].double(), steps=kw["sampling_steps"], shift=5.0,
                                guide_scale=kw["guide_scale"], cfg_star_switch=kw["cfg_star_switch"], cfg_zero_step=kw["cfg_zero_step"],
                                sample_solver=kw["sample_solver"], slg_layers=kw.get("slg_layers"), slg_start=kw.get("slg_start", 0.0),
                                slg_end=kw.get("slg_end", 1.0))
        assert W.rel_l2(lat, g["cases"][name]["latents"]) < 5e-5, name


def test_i2v_oracle_loop_and_product_mask_match_reference_generate_method(golden_dir):
    """WanI2V.generate, the reference's own method on a stand-in self (oracle/gen_golden_wan_generate.py:main_i2v): the oracle's i2v loop on
    `y = [mask | latent]` vs the returned latents — start image, and start + end image (one frame added, mask on both ends, last latent frame
    dropped) — and the PRODUCT's first_frame_mask (host-side torch) against the mask rows of that y, bit for bit."""
    from types import SimpleNamespace
    from ltx_video_gpupoor_b200.wan.image2video import WanI2V
    g = torch.load(os.path.join(golden_dir, "wan_i2v_generate.pt"), weights_only=False)
    cfg = g["cfg"]
    sd = {k: v.double() for k, v in W.make_wan_state_dict(cfg, seed=1).items()}
    pipe = WanI2V(SimpleNamespace(model_type="i2v"), device="cpu")
    assert len(g["cases"]) == 2
    for name, c in g["cases"].items():
        kw = c["kw"]
        assert torch.equal(pipe.first_frame_mask(c["frames"], 8, 12, c["any_end"], True), c["y"][:4]), name
        noise = torch.randn(16, c["lat_frames"], 8, 12, dtype=torch.float32, generator=torch.Generator().manual_seed(c["seed"]))
        with torch.no_grad():
            lat = W.t2v_denoise(sd, cfg, noise.double(), g["ctx"].double(), g["ctx0"].double(), steps=kw["sampling_steps"], shift=5.0,
                                guide_scale=kw["guide_scale"], cfg_star_switch=kw["cfg_star_switch"], cfg_zero_step=kw["cfg_zero_step"],
                                clip_fea=g["clip"].double(), y=c["y"].double())
        if c["any_end"]:
            lat = lat[:, :-1]
        assert W.rel_l2(lat, c["latents"]) < 5e-5, name


def test_seeded_wan_init_is_the_fixture_recipe():
    """bench.py's in-run sequence-parallel parity check builds its weights with the product-side `seeded_wan_state_dict`; they have to be
    the weights the reference fixture tests/golden/wan_t2v_h4.pt was recorded with (oracle.make_wan_state_dict), bit for bit."""
    import torch
    from ltx_video_gpupoor_b200.wan.init_weights import seeded_wan_state_dict
    from oracle import wan_oracle as W
    cfg = dict(W.WAN_1_3B, dim=512, ffn_dim=1280, num_heads=4, num_layers=2)
    a, b = W.make_wan_state_dict(cfg, seed=0), seeded_wan_state_dict(cfg, seed=0)
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)


def test_teacache_decision_is_taken_on_both_halves_under_cfg_parallel(monkeypatch):
    """under CFG-parallel the unconditional half only runs x_id = 1 forwards and used to follow a `should_calc` nobody
    updated.  With `_teacache_every_rank_decides` (set by WanT2V.generate when cfg_parallel is given) both halves take the same
    skip decisions from the replicated time embedding, and they are the decisions of the single-process x_id = 0 pass."""
    from ltx_video_gpupoor_b200 import ops
    from ltx_video_gpupoor_b200.wan.model import WanModel
    dist_seq = [0.0, 0.04, 0.03, 0.09, 0.02, 0.05, 0.2, 0.01]
    it = {}
    monkeypatch.setattr(ops, "rel_l1", lambda a, b: it["seq"].pop(0))

    def run(x_id, every):
        m = WanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=1)
        m.enable_teacache, m.rel_l1_thresh, m.teacache_start_step, m.num_steps = True, 0.1, 1, len(dist_seq)
        m._teacache_every_rank_decides = every
        it["seq"] = list(dist_seq)[1:] * 2
        return [m._teacache_should_calc(object(), i, x_id) for i in range(len(dist_seq))]

    cond = run(0, False)
    assert cond == run(0, True) == run(1, True)          # both CFG-parallel halves == the single-process decision
    assert False in cond and True in cond
    assert run(1, False) == [True] * len(dist_seq)       # the old behaviour: the uncond half never skipped


def _cond_par_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ltx_video_gpupoor_b200.ltx.distributed.cond_parallel import CondParallel
        cp = CondParallel(dist.group.WORLD)
        conds, bsz = 3, 2
        ranges = cp.ranges(conds)
        assert ranges == [(0, 2), (2, 3)] and cp.size == world and cp.rank == rank
        # every rank owns the rows of its conditions (value = 100 * cond + sample), the others hold garbage until the exchange
        pred = torch.full((conds * bsz, 4, 8), -1.0)
        lo, hi = ranges[rank]
        for c in range(lo, hi):
            for j in range(bsz):
                pred[c * bsz + j] = 100.0 * c + j
        cp.exchange(pred, bsz, ranges)
        want = torch.tensor([100.0 * c + j for c in range(conds) for j in range(bsz)])
        ret[rank] = (bool(torch.equal(pred[:, 0, 0], want)) and bool((pred == pred[:, :1, :1]).all()),
                     cp.any_flag(False, "cpu"), cp.any_flag(rank == 1, "cpu"))
    finally:
        dist.destroy_process_group()


def test_ltx_cond_parallel_exchange_gloo():
    """Guidance-condition parallelism of the LTX loop (ltx/distributed/cond_parallel.py) on CPU: 3 conditions over 2 ranks (2 + 1), every
    owner's prediction rows reach every rank; the `_interrupt` poll is an OR over the group."""
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_cond_par_worker, args=(world, 29500 + (os.getpid() % 2000) + 7, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert ret[r] == (True, False, True), (r, ret[r])
