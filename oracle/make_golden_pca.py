"""tests/golden/fn_pca_batch.npz: outputs of the UNMODIFIED reference utils.compute_pca_sklearn_batch (utils.py:392-470,
sklearn PCA per item) on seeded synthetic MC-dropout-like samples with a decaying spectrum.  Dev container only."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
GOLD = os.path.join(HERE, "..", "tests", "golden")
import ref_loader  # noqa: E402

ref_loader.install_shims(trainer=True)
import utils as ref_utils  # noqa: E402


sys.path.insert(0, os.path.join(HERE, "..", "tests"))
from helpers import pca_samples as samples  # noqa: E402


if __name__ == "__main__":
    pcs, scaled, weights, mean, svals = ref_utils.compute_pca_sklearn_batch(samples(), 5)
    np.savez_compressed(os.path.join(GOLD, "fn_pca_batch.npz"), pcs=pcs.numpy(), scaled=scaled.numpy(), weights=weights.numpy(),
                        mean=mean.numpy(), svals=svals.numpy())
    print("ok", pcs.shape)
