import sys, torch
sys.path.insert(0, '.')
import audioflux_b200 as af
c = af.CQT(84, 48000)
x = 0.1 * torch.randn((256, 240000), device='cuda')
for _ in range(2):
    c.cqt_batch(x)
torch.cuda.synchronize()
