"""Short isolated timing of regat_gemm for epilogue experiments (REGAT_TC_DBG=0|1|2).  Run on a GPU box."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.gemm_sweep import run  # noqa
for K in (256, 1024):
    for t in (1, 8):
        M = 148 * 128 * t
        us, tf = run(M, 256, K)
        print(f"M={M:6d} N=256 K={K:5d} bf16-out {us:8.1f} us {tf:8.1f} TF/s")
        us, tf = run(M, 256, K, c_f32=True)
        print(f"M={M:6d} N=256 K={K:5d} f32-out  {us:8.1f} us {tf:8.1f} TF/s")
for epi in ("plain", "bias", "acc"):
    us, tf = run(9216, 1024, 1024, epi_kind=epi)
    print(f"9216x1024x1024 {epi:6s} {us:8.1f} {tf:8.1f}")
us, tf = run(1024, 1024, 9216, tA=1, c_f32=True); print(f"wgrad 1024x1024x9216 {us:8.1f} {tf:8.1f}")
us, tf = run(1024, 2048, 9216, tA=1, c_f32=True); print(f"wgrad 1024x2048x9216 {us:8.1f} {tf:8.1f}")
us, tf = run(256, 1536, 768); print(f"small 256x1536x768 {us:8.1f} {tf:8.1f}")
us, tf = run(256, 768, 1024); print(f"small 256x768x1024 {us:8.1f} {tf:8.1f}")
us, tf = run(256, 3136, 1536, c_f32=True); print(f"small 256x3136x1536 {us:8.1f} {tf:8.1f}")
us, tf = run(5120, 4096, 1024); print(f"KV 5120x4096x1024 {us:8.1f} {tf:8.1f}")
us, tf = run(9216, 2048, 1024); print(f"Q 9216x2048x1024 {us:8.1f} {tf:8.1f}")
us, tf = run(9216, 1024, 2048); print(f"v2out 9216x1024x2048 {us:8.1f} {tf:8.1f}")
print("skinny set (M=256):")
for (M, N, K, tb) in ((256, 1536, 768, 0), (256, 1024, 768, 1), (256, 768, 1024, 0), (256, 1536, 3136, 1), (256, 768, 1536, 1), (256, 1024, 768, 1)):
    us, tf = run(M, N, K, tB=tb); print(f"  {M}x{N}x{K} tB={tb} {us:8.1f} us {tf:8.1f} TF/s")
us, tf = run(256, 3136, 1536, c_f32=True, epi_kind="bias"); print(f"  256x3136x1536 f32 bias {us:8.1f} {tf:8.1f}")
print("fixed-cost probes (graph replay):")
for (M, N, K) in ((128, 128, 64), (128, 128, 512), (128, 128, 4096), (256, 768, 64), (256, 768, 1024), (2048, 1024, 64), (18944, 256, 64)):
    us, tf = run(M, N, K); print(f"  {M}x{N}x{K} {us:8.1f} us")
