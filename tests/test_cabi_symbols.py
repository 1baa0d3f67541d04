"""The C-ABI library loads (no GPU needed) and exports every symbol include/cnsn_b200.h declares,
and the ctypes binding covers exactly that set.  No compute calls here."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cnsn_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cnsn_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("cnsn_instance_stats", "cnsn_selfnorm_fwd", "cnsn_selfnorm_bwd", "cnsn_crossnorm_fwd",
                 "cnsn_crossnorm_bwd", "cnsn_version", "cnsn_error_string", "cnsn_launch_count"):
        assert must in syms


def test_library_exports_every_declared_symbol(_built_library):
    lib = ctypes.CDLL(_built_library)
    for name in declared_symbols():
        assert hasattr(lib, name), "libcnsn_b200.so does not export %s" % name
    lib.cnsn_error_string.restype = ctypes.c_char_p
    assert lib.cnsn_error_string(0) == b"ok"
    assert b"more than 1 value per channel" in lib.cnsn_error_string(-3)


def test_binding_covers_exactly_the_header():
    import cnsn_b200._lib as L
    assert sorted(L.SIGNATURES) == declared_symbols()
    assert L.lib().cnsn_version() == L.ABI_VERSION
    src = open(HEADER).read()
    assert int(re.search(r"#define CNSN_ABI_VERSION (\d+)", src).group(1)) == L.ABI_VERSION


def test_argument_validation_without_gpu():
    """Bad arguments are rejected before any CUDA call, so this runs on a CPU-only box."""
    import cnsn_b200._lib as L
    h = L.lib()
    assert h.cnsn_instance_stats(None, 0, 1, 1, 1, 1, 0, 1, 0, 1, 1e-5, None, None, None) == -1
    assert h.cnsn_selfnorm_save_floats(4, 16, 0) == 4 * 64 + 16 + 2 * 64 + 36 * 16 + 8
    assert h.cnsn_selfnorm_save_floats(4, 16, 1) == 6 * 64 + 32 + 2 * 64 + 36 * 16 + 8
    assert h.cnsn_crossnorm_save_floats(4, 16) == 256 + 2 * 64 + 8
    # fused site: 8 statistics blocks of N*C, r (C), exchange area (sn words, channel sectors, ticket, 3 cn words)
    assert h.cnsn_site_workspace_floats(4, 16) == 8 * 64 + 8 * 16 + 8
    assert h.cnsn_site_save_floats(4, 16) == 8 * 64 + 16 + h.cnsn_site_workspace_floats(4, 16)
    assert h.cnsn_site_fwd(None, None, 0, 4, 16, 8, 8, None, None, None, 0.0, 1e-5, None, 0.1, 1e-5, 1e-12, 0, None, None) == -1
    assert h.cnsn_site_bwd(None, None, None, 0, 4, 16, 8, 8, None, None, None, 0.0, 1e-5, 0, None, None, None, None, None) == -1
    assert h.cnsn_site_supported(7, 4, 16, 8, 8) == 0 and h.cnsn_site_supported(0, 0, 16, 8, 8) == 0    # bad dtype / dims
    assert h.cnsn_site_supported(0, 4, 16, 7, 7) == 0          # 7x7 fp32 planes are not 16-byte multiples
    assert h.cnsn_ibn_fwd(None, None, 0, 4, 16, 8, 8, 8, None, 1, 0, 0.1, 1e-5, 1e-5, None, None) == -1


def test_error_strings_cover_every_code():
    import cnsn_b200._lib as L
    h = L.lib()
    for code in (0, -1, -2, -3, -4, -5):
        assert h.cnsn_error_string(code).decode() not in ("", "cnsn: unknown error"), code
    assert "unknown" in h.cnsn_error_string(-99).decode()
