"""Probe: what a plain streaming READ of the same 495 MB achieves on this GPU (torch reductions), as a
practical ceiling for the decode_compact kernel (MEASURED_PEAKS.json's figure is a read+write copy)."""
import torch

x = torch.randn(64 * 22743 * 85, device="cuda")
y = torch.empty_like(x)
for name, fn in (("sum", lambda: x.sum()), ("amax", lambda: x.amax()), ("copy", lambda: y.copy_(x)),
                 ("amax_dim", lambda: x.view(-1, 85 * 361).amax(1))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    nbytes = x.numel() * 4 * (2 if name == "copy" else 1)
    print(f"{name:10s} {ms * 1000:8.1f} us  {nbytes / ms / 1e6:8.1f} GB/s")
