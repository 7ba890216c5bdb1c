"""Checkpoint loading shared by the model classes (counterpart of reference src/network/base_model.py:4-16).

Two on-disk layouts exist in the reference's ecosystem: a bare state_dict (MiDaS release files, `torch.save(model.state_dict())`
in main.py) and a training checkpoint that wraps it as {"model": state_dict, "optimizer": ...}.  Both load here."""
import torch


def read_state_dict(path):
    """state_dict stored at `path`, unwrapping a training checkpoint (recognised by its "optimizer" entry)."""
    blob = torch.load(path, map_location="cpu")
    return blob["model"] if "optimizer" in blob else blob


class BaseModel(torch.nn.Module):
    def load(self, path):
        """`model.load(path)` as the reference's constructors call it (strict key matching)."""
        self.load_state_dict(read_state_dict(path))
