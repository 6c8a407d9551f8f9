import sys, torch
sys.path.insert(0, '.')
import audioflux_b200 as af
w = af.CWT(84, 19, 48000, wavelet_type=af.WaveletContinueType.MORLET, is_padding=False)
x = 0.1 * torch.randn((4, 1 << 19), device='cuda')
for _ in range(3):
    w.cwt_batch(x)
torch.cuda.synchronize()
