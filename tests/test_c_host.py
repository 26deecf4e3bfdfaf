"""include/itw_bcn.h is a C header (the reference's ispc_texcomp.h is consumed from C++ and could be from C): compile a small
C host with gcc -std=c99 -Wall -Werror, link it against libitw_bcn.so exactly as a maintainer would, and run it.  Without a
GPU the compute call must report an error through itw_get_last_error (no CPU fallback); with one it must succeed."""
import os
import subprocess
import tempfile

import itw_testlib as T

HOST = r"""
#include <stdio.h>
#include <string.h>
#include "itw_bcn.h"

int main(void)
{
    bc7_enc_settings s7;
    bc6h_enc_settings s6;
    memset(&s7, 0, sizeof s7);
    memset(&s6, 0, sizeof s6);
    GetProfile_slow(&s7);
    GetProfile_bc6h_slow(&s6);
    printf("sizes %zu %zu %zu\n", sizeof(rgba_surface), sizeof(bc7_enc_settings), sizeof(bc6h_enc_settings));
    printf("slow %d %d %d %d\n", s7.fastSkipTreshold_mode1, s7.fastSkipTreshold_mode3, s7.refineIterations[1], s7.channels);
    printf("bc6h %d %d %d\n", (int)s6.slow_mode, s6.fastSkipTreshold, s6.refineIterations_2p);
    unsigned char pixels[8 * 8 * 4], blocks[4 * 8];
    memset(pixels, 200, sizeof pixels);
    memset(blocks, 0xAB, sizeof blocks);
    rgba_surface surf = { pixels, 8, 8, 32 };
    CompressBlocksBC1(&surf, blocks);
    const char* err = itw_get_last_error();
    printf("error [%s] first byte %02X bpb %d\n", err, blocks[0], itw_bytes_per_block(ITW_FORMAT_BC1));
    itw_dds_desc d = { 64, 32, 3, 1, 98, 0 };
    printf("dds %zu %zu\n", itw_dds_header_bytes(&d), itw_dds_file_bytes(&d));
    return 0;
}
"""


def test_c99_host_compiles_links_and_runs():
    lib_dir = os.path.join(T.ROOT, "intel-texture-works-plugin_b200")
    T.product()                                                  # makes sure the library exists
    with tempfile.TemporaryDirectory() as tmp:
        src, exe = os.path.join(tmp, "host.c"), os.path.join(tmp, "host")
        open(src, "w").write(HOST)
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(T.ROOT, "include"), src, "-o", exe,
                               "-L", lib_dir, "-l:libitw_bcn.so", "-Wl,-rpath," + lib_dir])
        out = subprocess.check_output([exe], text=True).splitlines()
    assert out[0] == "sizes 24 64 16"
    assert out[1] == "slow 64 64 4 3"                            # ispc_texcomp.cpp:156-189
    assert out[2] == "bc6h 1 10 2"                               # ispc_texcomp.cpp:397-403
    has_gpu = False
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except ImportError:
        pass
    if has_gpu:
        assert out[3].startswith("error []") and "first byte AB" not in out[3], out[3]
    else:
        assert out[3].startswith("error [") and not out[3].startswith("error []") and "first byte AB" in out[3], out[3]
    assert out[3].endswith("bpb 8")
    assert out[4] == "dds 148 %d" % (148 + (16 * 8 + 8 * 4 + 4 * 2) * 16)
