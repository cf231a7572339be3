import torch
def try_capture(name, fn):
    torch.cuda.synchronize()
    try:
        fn(); fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        print(name, "OK", flush=True)
    except Exception as e:
        print(name, "FAILED", str(e).splitlines()[0], flush=True)
        torch.cuda.synchronize()
for B, M in ((2, 4320), (2, 17280), (4, 276480)):
    x = torch.rand(B, M, device="cuda")
    m = torch.rand(B, M, device="cuda") > 0.5
    try_capture("argsort %dx%d" % (B, M), lambda: x.argsort(dim=1))
    try_capture("sort %dx%d" % (B, M), lambda: torch.sort(x, dim=1))
    try_capture("topk(largest=False,k=M//5) %dx%d" % (B, M), lambda: torch.topk(x, M // 5, dim=1, largest=False))
    order = x.argsort(dim=1)
    try_capture("scatter_ %dx%d" % (B, M), lambda: torch.empty_like(order).scatter_(1, order, torch.arange(M, device="cuda").expand(B, M)))
    lab = torch.randint(0, 4, (B, M), device="cuda"); lab[:, ::7] = 3000
    cls = torch.randn(B, M, 4, device="cuda", requires_grad=True)
    def ce():
        cls.grad = None
        l = torch.nn.functional.cross_entropy(cls.reshape(-1, 4), lab.reshape(-1), reduction="none", ignore_index=3000)
        l.sum().backward()
    try_capture("cross_entropy none+ignore fwd/bwd %dx%d" % (B, M), ce)
