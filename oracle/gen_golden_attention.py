"""Pin oracle/attention_oracle.py against the UNMODIFIED reference `pay_attention` (utils/attention.py, imported under the refshim with
`offload.shared_state["_attention"] = "sdpa"`) and write tests/golden/pay_attention.pt.
Build container only (needs /root/reference):  python oracle/gen_golden_attention.py"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "refshim"))
import load_reference  # noqa: E402

load_reference.install()
from oracle import attention_oracle as A  # noqa: E402
from oracle.ltx_oracle import rel_l2  # noqa: E402

torch.set_grad_enabled(False)


def main():
    from utils.attention import pay_attention as ref_pay_attention, get_attention_modes, get_supported_attention_modes
    import wan.modules.attention as wan_attn
    assert open(os.path.join(load_reference.REFERENCE_ROOT, "utils", "attention.py")).read() == \
        open(os.path.join(load_reference.REFERENCE_ROOT, "wan", "modules", "attention.py")).read(), "the two entry points are one file"
    print("reference backends in this container:", get_attention_modes(), get_supported_attention_modes())
    g = torch.Generator().manual_seed(0)
    rn = lambda *s: torch.randn(*s, generator=g)
    cases = {}
    # a: plain fp32, cross-shaped
    cases["plain"] = dict(q=rn(2, 96, 3, 64), k=rn(2, 80, 3, 64), v=rn(2, 80, 3, 64), kw={})
    # b: additive key mask [B, 1, 1, Lk] (the LTX cross-attention mask, (1 - m) * -10000)
    m = torch.zeros(2, 1, 1, 80)
    m[0, ..., 50:] = -10000.0
    m[1, ..., 7:] = -10000.0
    cases["mask"] = dict(q=rn(2, 96, 3, 64), k=rn(2, 80, 3, 64), v=rn(2, 80, 3, 64), kw=dict(attention_mask=m))
    # c: batch of 3 with runs of equal key length (Wan joint pass with different prompt lengths)
    cases["k_lens_batch"] = dict(q=rn(3, 64, 2, 128), k=rn(3, 72, 2, 128), v=rn(3, 72, 2, 128), kw=dict(k_lens=torch.tensor([72, 40, 40])))
    # d: one sequence, padded queries and keys
    cases["q_k_lens_single"] = dict(q=rn(1, 100, 2, 64), k=rn(1, 90, 2, 64), v=rn(1, 90, 2, 64),
                                    kw=dict(q_lens=torch.tensor([77]), k_lens=torch.tensor([33])))
    # e: dtype contract: q fp32, k fp32, v bf16 -> computed in bf16, returned in fp32; softmax_scale is ignored on the sdpa path
    cases["dtypes"] = dict(q=rn(1, 48, 2, 64), k=rn(1, 48, 2, 64), v=rn(1, 48, 2, 64).bfloat16(), kw=dict(softmax_scale=123.0))
    out = {}
    for name, c in cases.items():
        for fn_name, fn in (("utils.attention", ref_pay_attention), ("wan.modules.attention", wan_attn.pay_attention)):
            lst = [c["q"].clone(), c["k"].clone(), c["v"].clone()]
            y_ref = fn(lst, **c["kw"])
            assert lst == [], "the reference empties the caller's list"
        kw = {k: v for k, v in c["kw"].items() if k != "softmax_scale"}
        lst = [c["q"].clone(), c["k"].clone(), c["v"].clone()]
        y = A.pay_attention(lst, **kw)
        assert lst == [] and y.dtype == y_ref.dtype == c["q"].dtype and y.shape == y_ref.shape
        valid = int(c["kw"]["q_lens"][0]) if "q_lens" in c["kw"] else y.shape[1]
        e = rel_l2(y[:, :valid].float(), y_ref[:, :valid].float())
        print(f"  pay_attention[{name}]: out {tuple(y.shape)} {y.dtype}, rel_l2(oracle, reference) = {e:.3e}")
        assert e < (1e-2 if name == "dtypes" else 2e-6)
        out[name] = dict(q=c["q"], k=c["k"], v=c["v"], kw=c["kw"], out=y_ref, valid=valid)
    torch.save(out, os.path.join(ROOT, "tests", "golden", "pay_attention.pt"))
    print("written tests/golden/pay_attention.pt")


if __name__ == "__main__":
    main()
