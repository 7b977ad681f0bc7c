"""Calibration: how fast can 54.85 MB canvases be written at all (cudaMemset / torch fill), ring of 12 (> L2)."""
import torch
bufs = [torch.empty((64, 496, 432), dtype=torch.float32, device="cuda") for _ in range(12)]
def t(fn, reps=60):
    for i in range(12): fn(bufs[i % 12])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(bufs[i % 12])
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps
print("zero_ (memset)   %.2f us" % t(lambda b: b.zero_()))
print("fill_(1.0)       %.2f us" % t(lambda b: b.fill_(1.0)))
one = bufs[0]
print("same buffer zero %.2f us" % t(lambda b: one.zero_()))
src = torch.empty_like(bufs[0])
print("copy_ 55MB       %.2f us" % t(lambda b: b.copy_(src)))
