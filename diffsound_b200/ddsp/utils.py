"""API mirror of src/ddsp/utils.py:6-9."""
import torch


def modifed_sigmoid(x):
    """2 * sigmoid(x)**2.3 + 1e-6 (the reference's spelling is kept: it is the public name)."""
    return 2 * (torch.sigmoid(x) ** 2.3) + 1e-6
