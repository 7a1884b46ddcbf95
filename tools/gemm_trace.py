"""Per-unit timeline of one tcgen05 GEMM launch (CTA 0): where a unit's time goes -- operand arrival, MMA issue, accumulator
completion, epilogue.  Run on a GPU box:  REGAT_TC_CTA2=0 python tools/gemm_trace.py 9216 2048 1024 [tA tB]"""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tf_vqa_regat_b200 import _lib

l = _lib.lib()
M, N, K = (int(x) for x in sys.argv[1:4])
tA, tB = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (0, 0)
f32 = len(sys.argv) > 6 and sys.argv[6] == "f32"
dev = torch.device("cuda")
A = torch.randn((K, M) if tA else (M, K), device=dev).to(torch.bfloat16)
B = torch.randn((N, K) if tB else (K, N), device=dev).to(torch.bfloat16)
Cm = torch.zeros(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
bias = torch.zeros(N, device=dev)
epi = _lib.Epilogue()
if not f32:
    epi.bias = bias.data_ptr(); epi.relu = 1
st = torch.cuda.current_stream().cuda_stream
call = lambda: _lib.check(l.regat_gemm(_lib.BF16, tA, tB, M, N, K, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), Cm.data_ptr(), N,
                                       _lib.F32 if f32 else _lib.BF16, C.byref(epi), st))
for _ in range(3):
    call()
buf = torch.zeros(256, device=dev, dtype=torch.int64)
l.regat_gemm_trace(buf.data_ptr())
call()
torch.cuda.synchronize()
l.regat_gemm_trace(None)
t = buf.cpu().tolist()
ghz = 1.965
t0 = t[0]
us = lambda x: (x - t0) / ghz / 1e3 if x else float("nan")
print(f"{M}x{N}x{K} tA={tA} tB={tB} CTA2={os.environ.get('REGAT_TC_CTA2', '1')}: roles done {us(t[1]):.2f} us, exit {us(t[2]):.2f} us (clock64 / {ghz} GHz, from kernel start)")
print("unit  tma_first tma_last | acc_free operands mma_issued | epi_ready acc_done stored   (us)")
for u in range(14):
    r = t[8 + 8 * u: 16 + 8 * u]
    if not any(r):
        break
    print(f"{u:4d}  {us(r[0]):8.2f} {us(r[1]):8.2f} | {us(r[2]):8.2f} {us(r[3]):8.2f} {us(r[4]):8.2f} | {us(r[5]):8.2f} {us(r[6]):8.2f} {us(r[7]):8.2f}")
