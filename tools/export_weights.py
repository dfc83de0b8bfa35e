"""Export the reference's deployed downwash MLP checkpoint to a plain .npz.

Source: /root/reference/ndp_nmpc/scripts/dnwash_nn_est/nn_model/
        128-64-128_WBias_SN=4_epoch=20000_test_loss=1.0221.pkl   (selected at downwash_nn.py:15)
The .pkl is a torch state_dict (keys 0/2/4/6 .weight/.bias, fp32, 17 859 parameters); the .npz
holds the same arrays so the GPU box (which has no /root/reference) can load them.
"""
import hashlib
import os

import numpy as np
import torch

REF = "/root/reference/ndp_nmpc/scripts/dnwash_nn_est/nn_model"
NAME = "128-64-128_WBias_SN=4_epoch=20000_test_loss=1.0221"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ndp_nmpc_qd_b200", "dnwash_nn_est", "nn_model", NAME + ".npz")

if __name__ == "__main__":
    p = os.path.join(REF, NAME + ".pkl")
    print("sha256", hashlib.sha256(open(p, "rb").read()).hexdigest())
    sd = torch.load(p, map_location="cpu", weights_only=True)
    np.savez(OUT, **{k: v.numpy() for k, v in sd.items()})
    print("wrote", os.path.normpath(OUT), {k: tuple(v.shape) for k, v in sd.items()})
