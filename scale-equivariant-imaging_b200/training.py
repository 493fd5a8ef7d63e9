"""Checkpoint I/O (reference: src/training.py:6-45).  The file format is the reference's -- one torch.save'd dict with the
keys epoch / params / optimizer / scheduler -- so checkpoints move between the two code bases; `params` is whatever
model.get_weights() returns (the fp32 masters here; the bf16 copies the GEMMs read are derived state and are rebuilt)."""
from pathlib import Path

import torch

STATE_KEYS = ("epoch", "params", "optimizer", "scheduler")
_HUB_URL = "https://huggingface.co/jscanvic/scale-equivariant-imaging/resolve/main/{name}.pt?download=true"


def training_state(epoch, model, optimizer, scheduler):
    """the checkpoint dict (reference :22-30)"""
    return dict(zip(STATE_KEYS, (epoch, model.get_weights(), optimizer.state_dict(), scheduler.state_dict())))


def save_training_state(epoch, model, optimizer, scheduler, state_path):
    Path(state_path).parent.mkdir(parents=True, exist_ok=True)
    print(f"writing the training state to the file {state_path}")
    torch.save(training_state(epoch, model, optimizer, scheduler), state_path)


def get_weights(weights_name, device):
    """weights from a local file if `weights_name` names one, otherwise the published checkpoint of that name (reference
    :33-45); a full training state is reduced to its `params` entry"""
    if Path(weights_name).exists():
        blob = torch.load(weights_name, map_location=device)
    else:
        blob = torch.hub.load_state_dict_from_url(_HUB_URL.format(name=weights_name), map_location=device)
    return blob["params"] if "params" in blob else blob
