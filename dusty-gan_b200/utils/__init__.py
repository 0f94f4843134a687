"""Value-range helpers on the path (reference utils/__init__.py:70-79, 213-214)."""
import torch


def sigmoid_to_tanh(x: torch.Tensor):
    """[0,1] -> [-1,+1]"""
    return x * 2.0 - 1.0


def tanh_to_sigmoid(x: torch.Tensor):
    """[-1,+1] -> [0,1]"""
    return (x + 1.0) / 2.0


def flatten(tensor_BCHW):
    """(B,C,H,W) -> (B,H*W,C) contiguous."""
    return tensor_BCHW.flatten(2).permute(0, 2, 1).contiguous()
