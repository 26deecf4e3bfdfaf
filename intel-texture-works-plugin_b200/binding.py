"""ctypes binding of include/itw_bcn.h (same names, argument meaning and error behaviour)."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def library_path():
    # ITW_BCN_LIB selects another build of the same sources (tools/tune_unroll.sh); there is no other implementation to fall back to
    return os.environ.get("ITW_BCN_LIB", os.path.join(HERE, "libitw_bcn.so"))


class RgbaSurface(ctypes.Structure):          # ispc_texcomp.h:19-25
    _fields_ = [("ptr", ctypes.c_void_p), ("width", ctypes.c_int32), ("height", ctypes.c_int32),
                ("stride", ctypes.c_int32)]


class Bc7Settings(ctypes.Structure):          # ispc_texcomp.h:27-41
    _fields_ = [("mode_selection", ctypes.c_bool * 4), ("refineIterations", ctypes.c_int * 8),
                ("skip_mode2", ctypes.c_bool), ("fastSkipTreshold_mode1", ctypes.c_int),
                ("fastSkipTreshold_mode3", ctypes.c_int), ("fastSkipTreshold_mode7", ctypes.c_int),
                ("mode45_channel0", ctypes.c_int), ("refineIterations_channel", ctypes.c_int),
                ("channels", ctypes.c_int)]


class Bc6hSettings(ctypes.Structure):         # ispc_texcomp.h:43-50
    _fields_ = [("slow_mode", ctypes.c_bool), ("fast_mode", ctypes.c_bool),
                ("refineIterations_1p", ctypes.c_int), ("refineIterations_2p", ctypes.c_int),
                ("fastSkipTreshold", ctypes.c_int)]


BC7_PROFILES = ("ultrafast", "veryfast", "fast", "basic", "slow",
                "alpha_ultrafast", "alpha_veryfast", "alpha_fast", "alpha_basic", "alpha_slow")
BC6H_PROFILES = ("bc6h_veryfast", "bc6h_fast", "bc6h_basic", "bc6h_slow", "bc6h_veryslow")

# name -> (DXGI format id, bytes per block, bytes per texel, settings kind)
FORMATS = {
    "BC1": (71, 8, 4, None), "BC3": (77, 16, 4, None), "BC4": (80, 8, 4, None), "BC5": (83, 16, 4, None),
    "BC6H": (95, 16, 8, "bc6h"), "BC7": (98, 16, 4, "bc7"),
}

EXPORTS = (["CompressBlocksBC1", "CompressBlocksBC3", "CompressBlocksBC4", "CompressBlocksBC5",
            "CompressBlocksBC6H", "CompressBlocksBC7", "itw_bytes_per_block", "itw_encode_device",
            "itw_encode_batch", "itw_set_device", "itw_get_last_error", "itw_kernel_launch_count",
            "itw_last_kernel_ms", "itw_dds_header_bytes", "itw_dds_image_bytes", "itw_dds_image_offset",
            "itw_dds_file_bytes", "itw_dds_write_header", "itw_dds_read_header", "itw_dds_encode_file",
            "itw_mip_scratch_bytes", "itw_generate_mips_device", "itw_dds_encode_texture", "itw_decode", "itw_convert_pixels", "itw_encode_pixels",
            "itw_generate_mips_device_f16", "itw_generate_mips_device_srgb", "itw_dds_encode_pixels", "itw_release", "itw_set_devices", "itw_get_devices",
            "itw_begin_deferred", "itw_flush", "GetProcessorCount", "InitWin32Threads", "DestroyThreads", "GetBytesPerBlock",
            "CompressImageMT", "CompressImageST", "CompressImageBC1", "CompressImageBC3",
            "itw_shard_plan_make", "itw_shard_unique_id", "itw_shard_init", "itw_shard_finalize", "itw_encode_mip_chain_sharded"]
           + ["CompressImageBC7_" + p for p in BC7_PROFILES] + ["CompressImage" + p.replace("bc6h_", "BC6H_") for p in BC6H_PROFILES]
           + ["GetProfile_" + p for p in BC7_PROFILES + BC6H_PROFILES])


class PixelSource(ctypes.Structure):          # include/itw_bcn.h section 6
    _fields_ = [("data", ctypes.c_void_p), ("width", ctypes.c_int32), ("height", ctypes.c_int32), ("planes", ctypes.c_int32),
                ("depth", ctypes.c_int32), ("row_bytes", ctypes.c_int64)]


FRONT_HAS_ALPHA, FRONT_GAMMA, FRONT_FLIP_X, FRONT_FLIP_Y, FRONT_NORMALIZE = 1, 2, 4, 8, 16


class DdsDesc(ctypes.Structure):              # include/itw_bcn.h section 3
    _fields_ = [("width", ctypes.c_uint32), ("height", ctypes.c_uint32), ("mip_levels", ctypes.c_uint32),
                ("array_size", ctypes.c_uint32), ("dxgi_format", ctypes.c_uint32), ("is_cubemap", ctypes.c_uint32)]


class ShardPlan(ctypes.Structure):            # include/itw_bcn.h section 7
    _fields_ = [("nranks", ctypes.c_int32), ("rank", ctypes.c_int32), ("levels", ctypes.c_int32), ("band_levels", ctypes.c_int32),
                ("band_y0", ctypes.c_int32), ("band_y1", ctypes.c_int32), ("slot_bytes", ctypes.c_uint64), ("chain_bytes", ctypes.c_uint64),
                ("level_offset", ctypes.c_uint64 * 16), ("level_bytes", ctypes.c_uint64 * 16), ("send_offset", ctypes.c_uint64 * 16),
                ("band_bytes", ctypes.c_uint64 * 16), ("texel_offset", ctypes.c_uint64), ("texel_bytes", ctypes.c_uint64)]


class EncoderApi:
    """A loaded library exporting the reference C-ABI (optionally with a symbol prefix).

    Used for the product (prefix "") and, by tests only, for the oracle ("oracle_"), the
    reference-source build ("") and the CPU emulation ("emu_")."""

    def __init__(self, path, prefix=""):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} is missing -- build it first (python __graft_entry__.py build)")
        self.path = path
        self.prefix = prefix
        self.lib = ctypes.CDLL(path)

    def fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def profile(self, name):
        """Settings struct filled by GetProfile_<name> (zero-initialised first, like the canonical model)."""
        s = Bc6hSettings() if name.startswith("bc6h_") else Bc7Settings()
        f = self.fn("GetProfile_" + name)
        f.restype = None
        f(ctypes.byref(s))
        return s

    def _call(self, fmt, surf, dst_ptr, settings):
        f = self.fn("CompressBlocks" + fmt)
        f.restype = None
        kind = FORMATS[fmt][3]
        if kind is None:
            f(ctypes.byref(surf), ctypes.c_void_p(dst_ptr))
        else:
            if settings is None:
                raise ValueError(f"{fmt} needs a settings struct")
            f(ctypes.byref(surf), ctypes.c_void_p(dst_ptr), ctypes.byref(settings))

    def encode(self, fmt, image, settings=None):
        """Encode a host numpy image (H x W x 4, uint8 for LDR / uint16 half bits for BC6H)."""
        _, bpb, texel, _ = FORMATS[fmt]
        h, w = image.shape[:2]
        assert image.strides[1] == texel and image.shape[2] * image.itemsize == texel, "RGBA8 / RGBA16F texels expected"
        surf = RgbaSurface(image.ctypes.data, w, h, image.strides[0])
        out = np.zeros((h // 4) * (w // 4) * bpb, dtype=np.uint8)
        self._call(fmt, surf, out.ctypes.data, settings)
        self.check()
        return out

    def decode(self, fmt, blocks, width, height):
        """itw_decode: host blocks -> H x W x 4 image (uint8; uint16 half bit patterns for BC6H)."""
        _, bpb, texel, _ = FORMATS[fmt]
        blocks = np.ascontiguousarray(np.frombuffer(bytes(blocks), np.uint8))
        assert blocks.size == (width // 4) * (height // 4) * bpb
        img = np.zeros((height, width, 4), np.uint16 if texel == 8 else np.uint8)
        self.decode_raw(fmt, blocks.ctypes.data, img.ctypes.data, width, height, img.strides[0])
        return img

    def decode_raw(self, fmt, blocks_ptr, dst_ptr, width, height, stride):
        """itw_decode on raw addresses (host or device)."""
        f = self.fn("itw_decode")
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(RgbaSurface)]
        surf = RgbaSurface(dst_ptr, width, height, stride)
        rc = f(FORMATS[fmt][0], ctypes.c_void_p(blocks_ptr), ctypes.byref(surf))
        self.check()
        if rc != 0:
            raise RuntimeError("itw_decode failed")

    def convert_pixels(self, fmt, pixels, flags=0, pad=True):
        """itw_convert_pixels: host numpy H x W x planes array (uint8 / uint16 / float32) -> RGBA8 / RGBA16F image."""
        h, w, planes = pixels.shape
        pixels = np.ascontiguousarray(pixels)
        texel = FORMATS[fmt][2]
        dw, dh = ((w + 3) & ~3, (h + 3) & ~3) if pad else (w, h)
        img = np.zeros((dh, dw, 4), np.uint16 if texel == 8 else np.uint8)
        src = PixelSource(pixels.ctypes.data, w, h, planes, pixels.itemsize * 8, 0)
        self.convert_pixels_raw(fmt, src, flags, img.ctypes.data, dw, dh, img.strides[0])
        return img

    def convert_pixels_raw(self, fmt, src, flags, dst_ptr, width, height, stride):
        f = self.fn("itw_convert_pixels")
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.c_int, ctypes.POINTER(PixelSource), ctypes.c_uint32, ctypes.POINTER(RgbaSurface)]
        surf = RgbaSurface(dst_ptr, width, height, stride)
        rc = f(FORMATS[fmt][0], ctypes.byref(src), flags, ctypes.byref(surf))
        self.check()
        if rc != 0:
            raise RuntimeError("itw_convert_pixels failed")

    def encode_raw(self, fmt, ptr, width, height, stride, dst_ptr, settings=None):
        """CompressBlocks<fmt> on raw addresses (host or device)."""
        self._call(fmt, RgbaSurface(ptr, width, height, stride), dst_ptr, settings)
        self.check()

    def check(self):
        pass


class ItwBcn(EncoderApi):
    """The product library.  Raises on any reported error; never falls back to a CPU path."""

    def __init__(self, path=None):
        super().__init__(path or library_path())
        L = self.lib
        L.itw_get_last_error.restype = ctypes.c_char_p
        L.itw_kernel_launch_count.restype = ctypes.c_uint64
        L.itw_last_kernel_ms.restype = ctypes.c_float
        L.itw_encode_device.restype = ctypes.c_int
        L.itw_encode_device.argtypes = [ctypes.c_int, ctypes.POINTER(RgbaSurface), ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_void_p]
        L.itw_encode_batch.restype = ctypes.c_int
        L.itw_bytes_per_block.restype = ctypes.c_int
        L.itw_set_device.restype = ctypes.c_int

        for n in ("itw_dds_header_bytes", "itw_dds_image_bytes", "itw_dds_image_offset", "itw_dds_file_bytes",
                  "itw_dds_write_header", "itw_dds_read_header", "itw_dds_encode_file"):
            getattr(L, n).restype = ctypes.c_size_t
        L.itw_dds_image_bytes.argtypes = [ctypes.POINTER(DdsDesc), ctypes.c_uint32]
        L.itw_dds_image_offset.argtypes = [ctypes.POINTER(DdsDesc), ctypes.c_uint32, ctypes.c_uint32]
        L.itw_dds_write_header.argtypes = [ctypes.POINTER(DdsDesc), ctypes.c_void_p, ctypes.c_size_t]
        L.itw_dds_read_header.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(DdsDesc)]
        L.itw_dds_encode_file.argtypes = [ctypes.POINTER(DdsDesc), ctypes.POINTER(RgbaSurface), ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_size_t]

        L.itw_mip_scratch_bytes.restype = ctypes.c_size_t
        L.itw_generate_mips_device.restype = ctypes.c_int
        L.itw_generate_mips_device.argtypes = [ctypes.POINTER(RgbaSurface), ctypes.c_int, ctypes.POINTER(RgbaSurface),
                                               ctypes.c_void_p, ctypes.c_void_p]
        L.itw_dds_encode_texture.restype = ctypes.c_size_t
        L.itw_dds_encode_texture.argtypes = [ctypes.POINTER(DdsDesc), ctypes.POINTER(RgbaSurface), ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_size_t]

    def encode_pixels(self, fmt, pixels, flags=0, settings=None):
        """itw_encode_pixels: host numpy H x W x planes array -> blocks of the padded image."""
        h, w, planes = pixels.shape
        pixels = np.ascontiguousarray(pixels)
        out = np.zeros(((w + 3) // 4) * ((h + 3) // 4) * FORMATS[fmt][1], np.uint8)
        src = PixelSource(pixels.ctypes.data, w, h, planes, pixels.itemsize * 8, 0)
        self.encode_pixels_raw(fmt, src, flags, out.ctypes.data, settings)
        return out

    def encode_pixels_raw(self, fmt, src, flags, dst_ptr, settings=None):
        f = self.lib.itw_encode_pixels
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.c_int, ctypes.POINTER(PixelSource), ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p]
        sp = ctypes.cast(ctypes.byref(settings), ctypes.c_void_p) if settings is not None else None
        rc = f(FORMATS[fmt][0], ctypes.byref(src), flags, sp, ctypes.c_void_p(dst_ptr))
        self.check()
        if rc != 0:
            raise RuntimeError("itw_encode_pixels failed")

    def last_error(self):
        return self.lib.itw_get_last_error().decode()

    def dds_encode_texture(self, desc, tops, settings=None):
        """itw_dds_encode_texture: `tops` = host numpy level-0 images (one per array item); mips made on the GPU."""
        n = self.lib.itw_dds_file_bytes(ctypes.byref(desc))
        if not n:
            raise ValueError("unsupported DDS description")
        surf = (RgbaSurface * len(tops))(*[RgbaSurface(im.ctypes.data, im.shape[1], im.shape[0], im.strides[0]) for im in tops])
        out = np.zeros(n, np.uint8)
        sp = ctypes.cast(ctypes.byref(settings), ctypes.c_void_p) if settings is not None else None
        got = self.lib.itw_dds_encode_texture(ctypes.byref(desc), surf, sp, out.ctypes.data, n)
        if got != n:
            self.check()
            raise RuntimeError("itw_dds_encode_texture failed")
        return out

    def dds_encode_pixels(self, desc, pixel_arrays, flags=0, settings=None):
        """itw_dds_encode_pixels: `pixel_arrays` = host numpy H x W x planes arrays (one per array item); the plug-in's whole save path."""
        n = self.lib.itw_dds_file_bytes(ctypes.byref(desc))
        if not n:
            raise ValueError("unsupported DDS description")
        keep = [np.ascontiguousarray(a) for a in pixel_arrays]
        srcs = (PixelSource * len(keep))(*[PixelSource(a.ctypes.data, a.shape[1], a.shape[0], a.shape[2], a.itemsize * 8, 0) for a in keep])
        out = np.zeros(n, np.uint8)
        f = self.lib.itw_dds_encode_pixels
        f.restype = ctypes.c_size_t
        f.argtypes = [ctypes.POINTER(DdsDesc), ctypes.POINTER(PixelSource), ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        sp = ctypes.cast(ctypes.byref(settings), ctypes.c_void_p) if settings is not None else None
        got = f(ctypes.byref(desc), srcs, flags, sp, out.ctypes.data, n)
        if got != n:
            self.check()
            raise RuntimeError("itw_dds_encode_pixels failed")
        return out

    def dds_encode_file(self, desc, images, settings=None):
        """itw_dds_encode_file: `images` = host numpy arrays, item-major / mip-minor, padded to multiples of 4."""
        n = self.lib.itw_dds_file_bytes(ctypes.byref(desc))
        if not n:
            raise ValueError("unsupported DDS description")
        surf = (RgbaSurface * len(images))(*[RgbaSurface(im.ctypes.data, im.shape[1], im.shape[0], im.strides[0]) for im in images])
        out = np.zeros(n, np.uint8)
        sp = ctypes.cast(ctypes.byref(settings), ctypes.c_void_p) if settings is not None else None
        got = self.lib.itw_dds_encode_file(ctypes.byref(desc), surf, sp, out.ctypes.data, n)
        if got != n:
            self.check()
            raise RuntimeError("itw_dds_encode_file failed")
        return out

    def check(self):
        e = self.last_error()
        if e:
            raise RuntimeError("libitw_bcn: " + e)

    def set_device(self, index):
        if self.lib.itw_set_device(int(index)) != 0:
            self.check()

    def set_devices(self, indices):
        """itw_set_devices: host -> host calls fan out over these GPUs (one process); [] = single-device behaviour."""
        arr = (ctypes.c_int * max(len(indices), 1))(*indices)
        self.lib.itw_set_devices.restype = ctypes.c_int
        self.lib.itw_set_devices.argtypes = [ctypes.POINTER(ctypes.c_int), ctypes.c_int]
        if self.lib.itw_set_devices(arr, len(indices)) != 0:
            self.check()
            raise RuntimeError("itw_set_devices failed")

    def begin_deferred(self):
        if self.lib.itw_begin_deferred() != 0:
            self.check()

    def flush(self):
        if self.lib.itw_flush() != 0:
            self.check()
            raise RuntimeError("itw_flush failed")

    def launch_count(self):
        return int(self.lib.itw_kernel_launch_count())

    def last_kernel_ms(self):
        return float(self.lib.itw_last_kernel_ms())

    def encode_batch(self, fmt, surfaces, dst_ptrs, settings=None):
        """itw_encode_batch: `surfaces` = list of (ptr, width, height, stride), `dst_ptrs` = list of addresses."""
        n = len(surfaces)
        arr = (RgbaSurface * n)(*[RgbaSurface(*s) for s in surfaces])
        dst = (ctypes.c_void_p * n)(*dst_ptrs)
        sp = ctypes.cast(ctypes.byref(settings), ctypes.c_void_p) if settings is not None else None
        self.lib.itw_encode_batch.argtypes = [ctypes.c_int, ctypes.POINTER(RgbaSurface), ctypes.POINTER(ctypes.c_void_p),
                                              ctypes.c_int, ctypes.c_void_p]
        if self.lib.itw_encode_batch(FORMATS[fmt][0], arr, dst, n, sp) != 0:
            self.check()
            raise RuntimeError("itw_encode_batch failed")

    def encode_device(self, fmt, ptr, width, height, stride, dst_ptr, settings=None, stream=0):
        """Enqueue the encode of a device-resident surface on a CUDA stream (no copies, no sync)."""
        surf = RgbaSurface(ptr, width, height, stride)
        sp = ctypes.cast(ctypes.byref(settings), ctypes.c_void_p) if settings is not None else None
        rc = self.lib.itw_encode_device(FORMATS[fmt][0], ctypes.byref(surf), ctypes.c_void_p(dst_ptr), sp,
                                        ctypes.c_void_p(stream))
        if rc != 0:
            self.check()
            raise RuntimeError("itw_encode_device failed")
