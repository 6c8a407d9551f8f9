#!/usr/bin/env python3
"""Compare the v2 fused kernel (mfcc_fused2.cu) with v1 (mfcc_fused.cu) and the numpy oracle on a few clips:
raw filter-bank output per filter / frame and the MFCC.  Prints where they differ."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
sys.path.insert(0, ROOT)
import torch
import audioflux_b200 as af
from oracle import af_oracle as O

def make(kernel):
    os.environ["AFB200_MFCC_KERNEL"] = kernel
    S, D = af.SpectralFilterBankScaleType, af.SpectralDataType
    return af.BFT(128, 11, 48000, slide_length=512, scale_type=S.MEL, data_type=D.POWER)

def main():
    B, Lc = int(os.environ.get("DBG_B", "3")), int(os.environ.get("DBG_L", "48000"))
    rng = np.random.default_rng(0)
    x = (0.1 * rng.standard_normal((B, Lc))).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    b2, b1 = make("v2"), make("v1")
    for name, fn in (("mel", lambda b: b.bft_batch(xd, result_type=1)), ("mfcc", lambda b: b.mfcc_batch(xd, 40))):
        r2 = fn(b2); torch.cuda.synchronize(); r2 = r2.cpu().numpy()
        r1 = fn(b1); torch.cuda.synchronize(); r1 = r1.cpu().numpy()
        want = np.stack([(O.bft(x[i], 128, 11, 48000, 512) if name == "mel" else O.mfcc(x[i], 48000, 11, 512, 128, 40)) for i in range(B)])
        s = np.abs(want).max()
        e2, e1 = np.abs(r2 - want) / s, np.abs(r1 - want) / s
        print(f"{name}: shape {r2.shape} v2 max rel err {e2.max():.3e}  v1 {e1.max():.3e}  nan v2 {np.isnan(r2).sum()}")
        if not e2.max() < 1e-4:
            bad = np.argwhere(~(e2 < 1e-4))
            print("  bad entries:", len(bad), "of", e2.size, " first:", bad[:10].tolist())
            print("  bad frames (mod 13):", np.bincount(bad[:, 1] % 13, minlength=13).tolist())
            print("  bad columns:", np.bincount(bad[:, 2], minlength=r2.shape[2]).tolist())
            f = bad[0]
            print("  v2 row", r2[f[0], f[1], :12], "\n  want  ", want[f[0], f[1], :12])
    print("DEBUG_DONE")

if __name__ == "__main__":
    main()
