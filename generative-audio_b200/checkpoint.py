"""Checkpoint format of the reference trainers (NPPCAudioTrainer.save_checkpoint, nppc_audio/trainer.py:319-335;
NPPCAudioInpaintingTrainer.save_checkpoint, inpainting/trainer/nppc_trainer.py:604-618): ONE file
`{'model_state_dict', 'optimizer_state_dict', 'step'}` whose model part the reference validators load strictly
(`validator.py:45-47`).  The reference has no resume path; `load_checkpoint` is the obvious inverse.
Derived caches (packed / folded / split 16-bit weights, LSTM plans) are never part of a state_dict: they are rebuilt from
the fp32 master parameters on the next forward (modules._NoDerivedState)."""
import os

import torch


class CheckpointMixin:
    """For the step classes (they hold `nppc_model` and `step`)."""

    def save_checkpoint(self, checkpoint_path: str, optimizer: torch.optim.Optimizer) -> str:
        checkpoint = {"model_state_dict": self.nppc_model.state_dict(), "optimizer_state_dict": optimizer.state_dict(),
                      "step": self.step}
        d = os.path.dirname(checkpoint_path)
        if d:
            os.makedirs(d, exist_ok=True)
        tmp = f"{checkpoint_path}.tmp{os.getpid()}"
        torch.save(checkpoint, tmp)
        os.replace(tmp, checkpoint_path)          # a crash mid-write never leaves a truncated checkpoint under the final name
        return checkpoint_path

    def load_checkpoint(self, checkpoint_path: str, optimizer: torch.optim.Optimizer = None, map_location=None) -> int:
        """Strict load of the model (as the reference validators do), optimizer state and step counter -> the restored step."""
        ck = torch.load(checkpoint_path, map_location=map_location or getattr(self.nppc_model, "device", None), weights_only=True)
        self.nppc_model.load_state_dict(ck["model_state_dict"], strict=True)
        if optimizer is not None and "optimizer_state_dict" in ck:
            optimizer.load_state_dict(ck["optimizer_state_dict"])
        self.step = int(ck.get("step", 0))
        if getattr(self, "_graph", None) is not None:
            self._graph = None                    # a captured training step holds the old optimizer state: capture again
        return self.step
