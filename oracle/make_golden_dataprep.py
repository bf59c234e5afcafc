"""tests/golden/fn_dataprep.npz from the UNMODIFIED reference dataset methods (AudioDataset._mix_with_snr / _normalize_audio,
AudioInpaintingDataset.time_to_spec_mask), called unbound on stub objects carrying only the config fields they read.
Dev container only."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
GOLD = os.path.join(HERE, "..", "tests", "golden")
import ref_loader  # noqa: E402

ref_loader.install_shims(trainer=True)
from dataset.audio_dataset import AudioDataset  # noqa: E402
from dataset.audio_dataset_inpainting import AudioInpaintingDataset  # noqa: E402

torch.set_grad_enabled(False)
rng = np.random.Generator(np.random.PCG64(99))
B, L = 4, 16000
clean = torch.from_numpy((rng.standard_normal((B, L)) * np.array([0.01, 0.2, 0.05, 0.6])[:, None]).astype(np.float32))
noise = torch.from_numpy((rng.standard_normal((B, L)) * np.array([0.3, 0.02, 0.05, 0.4])[:, None]).astype(np.float32))
snr = [-5.0, 0.0, 10.0, 20.0]
target = [-25.0, -25.0, -20.0, -3.0]     # the last row clips (peak > 0.99)
noisy_out, clean_out = [], []
for b in range(B):
    stub = types.SimpleNamespace(config=types.SimpleNamespace(target_dB_FS=target[b], target_dB_FS_floating_value=0.0))
    stub._normalize_audio = types.MethodType(AudioDataset._normalize_audio, stub)
    n, c = AudioDataset._mix_with_snr(stub, clean[b:b + 1], noise[b:b + 1], snr[b])
    noisy_out.append(n)
    clean_out.append(c)
noisy_out, clean_out = torch.stack(noisy_out), torch.stack(clean_out)
print("peaks", noisy_out.abs().amax(dim=1))

Lm, T_frames, win, hop = 64000, 500, 255, 128
masks = torch.ones(3, Lm)
masks[0, 20000:22304] = 0
masks[1, 0:300] = 0
masks[2, 63000:] = 0
stub2 = types.SimpleNamespace(config=types.SimpleNamespace(stft_configuration=types.SimpleNamespace(win_length=win, hop_length=hop)))
spec_c = torch.stack([AudioInpaintingDataset.time_to_spec_mask(stub2, masks[b:b + 1], T_frames, Lm, center=True) for b in range(3)])
spec_nc = torch.stack([AudioInpaintingDataset.time_to_spec_mask(stub2, masks[b:b + 1], T_frames, Lm, center=False) for b in range(3)])
np.savez_compressed(os.path.join(GOLD, "fn_dataprep.npz"), clean=clean.numpy(), noise=noise.numpy(), snr=np.array(snr, dtype=np.float32),
                    target=np.array(target, dtype=np.float32), noisy_out=noisy_out.numpy(), clean_out=clean_out.numpy(),
                    mask_gaps=np.array([[20000, 22304], [0, 300], [63000, Lm]]), mask_len=np.array([Lm]),
                    stft=np.array([T_frames, win, hop]), spec_center=spec_c.numpy(), spec_nocenter=spec_nc.numpy())
print("ok", spec_c.sum(dim=1), spec_nc.sum(dim=1))
