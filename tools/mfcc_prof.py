import sys, torch
sys.path.insert(0, '.')
import audioflux_b200 as af
S, D = af.SpectralFilterBankScaleType, af.SpectralDataType
b = af.BFT(128, 11, 48000, slide_length=512, scale_type=S.MEL, data_type=D.POWER)
x = 0.1 * torch.randn((1024, 240000), device='cuda')
for _ in range(3):
    b.mfcc_batch(x, 40)
torch.cuda.synchronize()
