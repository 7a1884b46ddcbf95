"""Isolated timing sweep of regat_gemm (tcgen05 path): fixed cost per launch vs per-tile cost.  Run on a GPU box."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tf_vqa_regat_b200 import _lib

l = _lib.lib()
dev = torch.device("cuda")
st = lambda: torch.cuda.current_stream().cuda_stream


def run(M, N, K, tA=0, tB=0, epi_kind="plain", c_f32=False, reps=20, nrot=4):
    dt = torch.bfloat16
    As = [torch.randn((K, M) if tA else (M, K), device=dev).to(dt) for _ in range(nrot)]
    Bm = torch.randn((N, K) if tB else (K, N), device=dev).to(dt)
    Cs = [torch.zeros(M, N, device=dev, dtype=torch.float32 if c_f32 else dt) for _ in range(nrot)]
    bias = torch.zeros(N, device=dev)
    epi = _lib.Epilogue()
    if epi_kind in ("bias", "acc"):
        epi.bias = bias.data_ptr(); epi.relu = 1
    if epi_kind == "acc":
        epi.accumulate = 1
    cd = _lib.F32 if c_f32 else _lib.BF16
    call = lambda i: _lib.check(l.regat_gemm(_lib.BF16, tA, tB, M, N, K, As[i % nrot].data_ptr(), As[0].stride(0), Bm.data_ptr(), Bm.stride(0),
                                             Cs[i % nrot].data_ptr(), N, cd, C.byref(epi), st()))
    for i in range(4):
        call(i)
    torch.cuda.synchronize()
    # replay from a CUDA graph: the Python / ctypes launch path (~10 us per call on these hosts) must not be what is timed
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(reps):
                call(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    return us, 2.0 * M * N * K / us / 1e6


if __name__ == "__main__":
    print("shape                         epi    us      TF/s   tiles/CTA")
    for K in (256, 1024, 2048, 4096):
        for t in (1, 2, 4, 8):
            M = 148 * 128 * t
            us, tf = run(M, 256, K)
            print(f"M={M:6d} N=256 K={K:5d} plain {us:8.1f} {tf:8.1f}   {t}")
    for epi in ("plain", "bias", "acc"):
        us, tf = run(9216, 1024, 1024, epi_kind=epi)
        print(f"9216x1024x1024 {epi:6s} {us:8.1f} {tf:8.1f}")
    us, tf = run(9216, 1024, 1024, tB=1); print(f"9216x1024x1024 dgrad(tB) {us:8.1f} {tf:8.1f}")
    us, tf = run(1024, 1024, 9216, tA=1, c_f32=True); print(f"wgrad 1024x1024x9216 {us:8.1f} {tf:8.1f}")
    us, tf = run(1024, 2048, 9216, tA=1, c_f32=True); print(f"wgrad 1024x2048x9216 {us:8.1f} {tf:8.1f}")
    us, tf = run(256, 1536, 768); print(f"small 256x1536x768 {us:8.1f} {tf:8.1f}")
    us, tf = run(256, 768, 1024); print(f"small 256x768x1024 {us:8.1f} {tf:8.1f}")
    # empty-ish kernel launch floor
    x = torch.zeros(1024, device=dev)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50): x.add_(1)
    e1.record(); torch.cuda.synchronize(); print("tiny torch kernel us", e0.elapsed_time(e1) / 50 * 1e3)
