"""Precision study (CPU): how far do reduced-precision tensor-core operand formats move the
CAMs away from the fp32 oracle?  Emulates rounding of BOTH operands of every matmul in the
ViT (fp32 accumulation), which is what tcgen05 kind::f16 / kind::tf32 do.
Run:  python tools/precision_study.py [size]
"""
import sys
sys.path.insert(0, ".")
import torch
from oracle import dupl_oracle as O


def q_fp16(t): return t.to(torch.float16).to(torch.float32)
def q_bf16(t): return t.to(torch.bfloat16).to(torch.float32)
def q_tf32(t):  # round-to-nearest-even to 10 mantissa bits
    i = t.contiguous().view(torch.int32)
    r = ((i >> 13) & 1) + 0x0FFF
    return ((i + r) & ~0x1FFF).view(torch.float32)
def q_tf32_trunc(t):
    i = t.contiguous().view(torch.int32)
    return (i & ~0x1FFF).view(torch.float32)
def q_fp16x2(t):  # hi+lo split: ~22 bits
    hi = t.to(torch.float16).to(torch.float32)
    lo = (t - hi).to(torch.float16).to(torch.float32)
    return hi + lo


def init_params(K=20, seed=0):
    """Random-init parameters with the reference's initialisation scheme (vit.py:262-275)."""
    g = torch.Generator().manual_seed(seed)
    P = {}
    def tn(*s): return torch.nn.init.trunc_normal_(torch.empty(*s), std=0.02, a=-2, b=2, generator=g)
    for br in (1, 2):
        e = f"branch{br}.encoder."
        P[e + "cls_token"] = tn(1, 1, 768); P[e + "pos_embed"] = tn(1, 197, 768)
        P[e + "patch_embed.proj.weight"] = (torch.rand(768, 3, 16, 16, generator=g) - 0.5) * 2 / (768 ** 0.5)
        P[e + "patch_embed.proj.bias"] = (torch.rand(768, generator=g) - 0.5) * 2 / (768 ** 0.5)
        for i in range(12):
            b = e + f"blocks.{i}."
            for n, (o, k) in {"attn.qkv": (2304, 768), "attn.proj": (768, 768), "mlp.fc1": (3072, 768), "mlp.fc2": (768, 3072)}.items():
                P[b + n + ".weight"] = tn(o, k); P[b + n + ".bias"] = torch.zeros(o)
            for n in ("norm1", "norm2"):
                P[b + n + ".weight"] = torch.ones(768); P[b + n + ".bias"] = torch.zeros(768)
        P[e + "norm.weight"] = torch.ones(768); P[e + "norm.bias"] = torch.zeros(768)
        for n in ("classifier", "aux_classifier"):
            P[f"branch{br}.{n}.weight"] = (torch.rand(K, 768, 1, 1, generator=g) - 0.5) * 2 / (768 ** 0.5)
    return P


if __name__ == "__main__":
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 224
    torch.manual_seed(0)
    P = init_params()
    x = torch.randn(1, 3, S, S)
    scales = (1.0, 0.5, 1.5)
    with torch.no_grad():
        ref_cam, ref_aux = O.multi_scale_cam(P, 1, x, scales)
        ca0, c0 = O.network_cam_only(P, 1, x)
        for name, q in [("tf32_rn", q_tf32), ("tf32_trunc", q_tf32_trunc), ("fp16", q_fp16), ("bf16", q_bf16), ("fp16x2", q_fp16x2)]:
            cam, aux = O.multi_scale_cam(P, 1, x, scales, quant=q)
            ca, c = O.network_cam_only(P, 1, x, quant=q)
            rel = lambda a, b: ((a - b).abs().max() / b.abs().max()).item()
            rms = lambda a, b: ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()
            print(f"{name:11s} raw cam max-rel {rel(c, c0):.2e} rms-rel {rms(c, c0):.2e} | aux {rel(ca, ca0):.2e} | "
                  f"ms-cam max-abs {((cam - ref_cam).abs().max()).item():.2e} aux {((aux - ref_aux).abs().max()).item():.2e}")
