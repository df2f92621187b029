import sys, torch
sys.path.insert(0, '/root/repo')
from leg_slam_b200 import cosine_query
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(3)
for P, Q in ((1000, 5), (4133, 256), (300, 300), (128, 16), (70000, 64)):
    f = (torch.randn(P, 64, generator=g) * (0.2 + torch.rand(P, 1, generator=g))).to(dev)
    t = torch.randn(Q, 64, generator=g).to(dev)
    ref = torch.nn.functional.normalize(f.double(), dim=1) @ torch.nn.functional.normalize(t.double(), dim=1).t()
    a = cosine_query(f, t); torch.cuda.synchronize()
    b = cosine_query(f, t, simt=True); torch.cuda.synchronize()
    print(P, Q, "tc err", float((a.double() - ref).abs().max()), "simt err", float((b.double() - ref).abs().max()), flush=True)
f = torch.randn(2_000_000, 64, generator=g).to(dev); t = torch.randn(256, 64, generator=g).to(dev)
for name, kw in (("tc", {}), ("simt", dict(simt=True))):
    for _ in range(3): cosine_query(f, t, **kw)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(10): cosine_query(f, t, **kw)
    e1.record(); torch.cuda.synchronize(); print(name, "cfgE ms", e0.elapsed_time(e1) / 10, flush=True)
x = torch.empty(2_000_000 * 256, device=dev)
for _ in range(3): x.fill_(1.0)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
for _ in range(10): x.fill_(1.0)
e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 10
print("fill 2GB ms", ms, "GB/s", x.numel() * 4 / ms / 1e6, flush=True)
