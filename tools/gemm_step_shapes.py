"""Isolated, CUDA-graph-replayed timing of the train step's large dense products (regat_gemm, tcgen05 path).
Run on a GPU box; REGAT_TC_CTA2=0/1 selects single CTAs / CTA pairs for the 256-wide tiles:
    for v in 0 1; do REGAT_TC_CTA2=$v python tools/gemm_step_shapes.py; done"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm_sweep import run

SHAPES = [  # name, M, N, K, tA, tB, epilogue, fp32 output
    ("v2out fwd", 9216, 1024, 2048, 0, 0, "bias", False),
    ("s fwd", 9216, 1024, 1024, 0, 0, "bias", False),
    ("Q fwd", 9216, 2048, 1024, 0, 0, "bias", False),
    ("KV fwd", 5120, 4096, 1024, 0, 0, "bias", False),
    ("dQ dgrad", 9216, 1024, 2048, 0, 1, "acc", False),
    ("dKV dgrad", 5120, 1024, 4096, 0, 1, "plain", False),
    ("dv0 dgrad", 9216, 1024, 1024, 0, 1, "acc", False),
    ("wgrad Q", 1024, 2048, 9216, 1, 0, "plain", True),
    ("wgrad KV", 1024, 4096, 5120, 1, 0, "plain", True),
    ("wgrad v2out", 2048, 1024, 9216, 1, 0, "plain", True),
    ("wgrad self", 1024, 1024, 9216, 1, 0, "plain", True),
    ("logits", 256, 3136, 1536, 0, 0, "bias", True),
    ("pv", 256, 768, 1024, 0, 0, "bias", False),
]
if __name__ == "__main__":  # noqa
    print("REGAT_TC_CTA2 =", os.environ.get("REGAT_TC_CTA2", "(default 1)"))
    tot = 0.0
    for name, M, N, K, tA, tB, epi, f32 in SHAPES:
        us, tf = run(M, N, K, tA=tA, tB=tB, epi_kind=epi, c_f32=f32)
        tot += us
        print(f"{name:12s} {M:5d}x{N:4d}x{K:4d} {epi:5s} {us:8.1f} us {tf:8.1f} TF/s")
    print(f"sum {tot:.1f} us")
