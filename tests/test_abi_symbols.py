"""CPU: the C-ABI library loads without a GPU and exports every symbol include/fr_b200.h declares; entry points that
need a device fail loudly (FR_ENODEVICE) instead of falling back to the CPU."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "fr_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fr_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_something():
    syms = declared_symbols()
    assert "fr_gallery_topk" in syms and "fr_gallery_sims" in syms and len(syms) >= 10


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(str(built_lib))
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/fr_b200.h but not exported: {missing}"


def test_abi_version_and_error_string(built_lib):
    lib = ctypes.CDLL(str(built_lib))
    lib.fr_last_error.restype = ctypes.c_char_p
    assert lib.fr_abi_version() >= 1
    assert isinstance(lib.fr_last_error(), bytes)


def test_no_cpu_fallback_without_device(built_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import frb200

    with pytest.raises(frb200.FrError) as e:
        frb200.Gallery.synthetic(16, seed=1)
    assert e.value.code == frb200.FR_ENODEVICE
    assert "no CPU fallback" in e.value.msg
    with pytest.raises(frb200.FrError) as e:    # the JPEG decoder too: no libjpeg path behind it
        frb200.JpegDecoder()
    assert e.value.code == frb200.FR_ENODEVICE


def test_header_is_plain_c(tmp_path):
    """include/fr_b200.h is the FFI boundary: it must compile as strict C99 (no C++-isms, every type it uses declared)"""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    src = tmp_path / "hdr.c"
    src.write_text('#include "fr_b200.h"\nint main(void) { FrJpegDecoder *d = 0; FrGallery *g = 0; (void)d; (void)g; return fr_abi_version() < 0; }\n')
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", f"-I{ROOT / 'include'}", str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr

