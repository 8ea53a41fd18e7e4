"""dev: fp32 GEMM options for the node-MLP on B200 (cuBLAS fp32 SIMT vs TF32 vs BF16x9 emulation if available)."""
import os, sys, time, torch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2560
a = torch.randn(N, 1000, device="cuda"); w = torch.randn(1000, 1000, device="cuda") * 0.03
ref = (a.double() @ w.double())
def bench(tag):
    for _ in range(5): (a @ w)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(50): c = a @ w
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    err = ((c.double() - ref).abs().max() / ref.abs().max()).item()
    print(f"{tag:28s} {ms*1e3:8.1f} us  {2*N*1e6/ms/1e9:8.1f} TFLOP/s  max rel err {err:.2e}")
torch.backends.cuda.matmul.allow_tf32 = False; bench("fp32 (allow_tf32=False)")
torch.backends.cuda.matmul.allow_tf32 = True; bench("tf32")
torch.backends.cuda.matmul.allow_tf32 = False
print("env CUBLAS_EMULATE_SINGLE_PRECISION =", os.environ.get("CUBLAS_EMULATE_SINGLE_PRECISION"))
ah, al = a.bfloat16(), (a - a.bfloat16().float()).bfloat16()
wh, wl = w.bfloat16(), (w - w.bfloat16().float()).bfloat16()
def bf16x3():
    return (ah @ wh).float() + (ah @ wl).float() + (al @ wh).float()
for _ in range(3): bf16x3()
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
for _ in range(50): c = bf16x3()
e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 50
print(f"{'bf16x3 (3 cuBLAS GEMMs)':28s} {ms*1e3:8.1f} us  max rel err {((c.double()-ref).abs().max()/ref.abs().max()).item():.2e}")
