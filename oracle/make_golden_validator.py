"""tests/golden/fn_pc_variations.npz from the UNMODIFIED reference functions used by NPPCAudioValidator
(decompress_cIRM, utils.crm_to_spectogram, torch.istft + the peak normalisation of save_audio_files).  Dev container only."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
GOLD = os.path.join(HERE, "..", "tests", "golden")
import ref_loader  # noqa: E402

ns = ref_loader.load_reference(trainer=False)
import utils as ref_utils  # noqa: E402

torch.set_grad_enabled(False)
rng = np.random.Generator(np.random.PCG64(77))
B, n, L = 2, 3, 4096
noisy = torch.from_numpy((rng.standard_normal((B, L)) * 0.05).astype(np.float32))
win = torch.hann_window(512)
noisy_c = torch.stft(noisy, 512, hop_length=256, win_length=512, window=win, return_complex=True)   # validator.py:72-79
Fq, T = noisy_c.shape[1:]
w_mat = torch.from_numpy((rng.standard_normal((B, n, 2, Fq, T)) * 4.0).astype(np.float32))   # |m| > 9.9 occurs: clipping path
enh_c = torch.stft(noisy * 0.7, 512, hop_length=256, win_length=512, window=win, return_complex=True)
alphas = torch.linspace(-3, 3, 6)
pcs, waves = [], []
for d in range(n):
    crm = ns.decompress_cIRM(w_mat[:, d]).permute(0, 2, 3, 1)            # validator.py:89-91
    pc = ref_utils.crm_to_spectogram(crm, noisy_c)                        # validator.py:94
    pcs.append(pc)
    row = []
    for a in alphas:
        var = enh_c + a * pc                                              # validator.py:266
        wv = torch.istft(var, 512, hop_length=256, win_length=512, window=win, length=L)   # :275-282
        row.append(wv / (wv.abs().amax(dim=-1, keepdim=True) + 1e-8))
    waves.append(torch.stack(row, dim=1))
pc = torch.stack(pcs, dim=1)
np.savez_compressed(os.path.join(GOLD, "fn_pc_variations.npz"), w_mat=w_mat.numpy(), noisy_real=noisy_c.real.numpy(),
                    noisy_imag=noisy_c.imag.numpy(), enh_real=enh_c.real.numpy(), enh_imag=enh_c.imag.numpy(),
                    alphas=alphas.numpy(), pc_real=pc.real.numpy(), pc_imag=pc.imag.numpy(),
                    variations=torch.stack(waves, dim=1).numpy(), length=np.array([L]))
print("ok", pc.shape, torch.stack(waves, dim=1).shape)
