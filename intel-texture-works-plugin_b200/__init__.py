"""itw-bcn-b200: B200-native BCn block encoder behind the Intel ISPC Texture Compressor C-ABI.

The product is `libitw_bcn.so` (csrc/, built by build.py).  This Python package is only the thin
ctypes binding the tests and bench use; it mirrors the reference's C interface
(3rdParty/Intel/Source/ispc_texcomp.h) one to one.  There is no CPU fallback: loading fails
loudly if the library is missing, and every encode raises if the CUDA path reports an error.
"""
from .binding import (  # noqa: F401
    BC6H_PROFILES,
    BC7_PROFILES,
    FORMATS,
    Bc6hSettings,
    Bc7Settings,
    DdsDesc,
    ItwBcn,
    RgbaSurface,
    library_path,
)
from . import synth  # noqa: F401
