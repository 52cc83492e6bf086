import torch, torch.nn.functional as F
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = True
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
for (N, C, Fd, T, dil) in ((8, 64, 64, 2048, 1), (8, 96, 128, 1024, 2), (8, 128, 256, 256, 4), (8, 256, 448, 32, 8)):
    x = torch.randn(N, C, Fd, T, device="cuda")
    w = torch.randn(C, C, 5, 3, device="cuda") * 0.02
    pad = (dil * 2, 1)
    xc, wc = x.contiguous(memory_format=torch.channels_last), w.contiguous(memory_format=torch.channels_last)
    a = t(lambda: F.conv2d(x, w, None, 1, pad, (dil, 1)))
    b = t(lambda: F.conv2d(xc, wc, None, 1, pad, (dil, 1)))
    c = t(lambda: torch.nn.grad.conv2d_input(x.shape, w, x, 1, pad, (dil, 1)))
    d = t(lambda: torch.nn.grad.conv2d_input(x.shape, wc, xc, 1, pad, (dil, 1)))
    cp = t(lambda: x.clone())
    print(f"{(N,C,Fd,T,dil)}: fprop nchw {a:.3f} ms, nhwc {b:.3f} ms | dgrad nchw {c:.3f}, nhwc {d:.3f} | copy {cp:.3f}")
