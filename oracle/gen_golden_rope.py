"""Wan RoPE tables with and without RIFLEx: the product's host function against the UNMODIFIED reference
(wan/modules/posemb_layers.py:432-473, 8-62) — bit-exact — and a subsampled fixture tests/golden/wan_rope_riflex.pt.
Build container only (needs /root/reference):  python oracle/gen_golden_rope.py"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "refshim"))
import load_reference  # noqa: E402

load_reference.install()


def main():
    from wan.modules.posemb_layers import get_rotary_pos_embed as ref
    from ltx_video_gpupoor_b200.wan.posemb_layers import get_rotary_pos_embed as ours
    out = {}
    for size in ((21, 60, 104), (33, 8, 12), (5, 4, 6)):
        for rf in (False, True):
            a, b = ref(size, enable_RIFLEx=rf), ours(size, enable_RIFLEx=rf)
            assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]), (size, rf)
            if size != (21, 60, 104):
                out[(size, rf)] = (a[0][::7].clone(), a[1][::7].clone())
    torch.save(out, os.path.join(ROOT, "tests", "golden", "wan_rope_riflex.pt"))
    print("RoPE tables (plain + RIFLEx) bit-exact vs the reference; written tests/golden/wan_rope_riflex.pt")


if __name__ == "__main__":
    main()
