import sys, torch
sys.path.insert(0, '/root/repo')
from leg_slam_b200 import cosine_query
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(3)
f = torch.randn(2_000_000, 64, generator=g).to(dev); t = torch.randn(256, 64, generator=g).to(dev)
for _ in range(3): cosine_query(f, t)
torch.cuda.synchronize(); print("ok")
