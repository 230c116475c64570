"""The reference's `wfd.wfdf` name that is on the guidance path (wfd/wfdf/losses.py:43-75)."""
from .losses import WFCLossEinsum

__all__ = ["WFCLossEinsum"]
