"""blockcopy -- B200-native drop-in for the reference ``blockcopy`` package.

Public names are the reference's (blockcopy/__init__.py:1-4).  All block movement runs in
hand-written sm_100a CUDA behind a C ABI (``blockcopy._C`` -> libblockcopy_sm100.so); there is no
CPU, CuPy or Triton path.
"""
from blockcopy.core.tensorwrapper import TensorWrapper, is_block, is_tensorwrapper, to_tensorwrapper, to_tensor
from blockcopy.core.blockcopy import BlockCopyModel, blockcopy_noblocks
from blockcopy.core.argparser import add_argparser_arguments
from blockcopy.policy.policy import build_policy_from_settings
from blockcopy.core.frame import U8Frame  # not in the reference: a decoded uint8 frame as lazily normalised input

__all__ = [
    "TensorWrapper", "is_block", "is_tensorwrapper", "to_tensorwrapper", "to_tensor",
    "BlockCopyModel", "blockcopy_noblocks", "add_argparser_arguments", "build_policy_from_settings",
]
