"""Per-stage diagnosis of the device input pipeline against the numpy checker (resized image, threshold, mask)."""
import os, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from oracle import input_oracle as IO
from causal_vae_b200.vessel.dataset import VesselBatchTransform
shapes = [(70, 90, 33, 35), (128, 64, 17, 130), (31, 29, 64, 200), (900, 40, 20, 64), (512, 512, 256, 256), (300, 420, 128, 96)]
for hin, win, H, W in shapes:
    raws = np.stack([IO.raw_image(hin, win, 20 + i) for i in range(3)])
    aug = [3, 0, 1]
    tf = VesselBatchTransform(H, W, 19)
    x, thr = tf.transform(torch.from_numpy(raws).cuda(), torch.tensor(aug), return_threshold=True)
    res = tf._ws[(3, x.device)][0].cpu().numpy()
    x = x.cpu().numpy()
    for i in range(3):
        o = IO.flip(IO.resize_aa(raws[i], H, W), aug[i])
        mask, othr, band = IO.preprocess_image(raws[i], H, W, aug[i])
        bad = np.argwhere(res[i] != o)
        print((hin, win, H, W), "img", i, "resized mismatches", len(bad), "first", bad[:3].tolist(),
              "thr", float(thr[i]), float(othr), "mask mismatches", int((x[i, 0] != mask[0]).sum()), "band", int(band.sum()))
