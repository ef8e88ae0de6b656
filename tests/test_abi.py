"""The C-ABI library loads without a GPU and exports every symbol include/vlmclip.h declares, with the argument
counts the ctypes binding uses.  No compute calls here."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "vlmclip.h").read_text()


def _declarations():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    decls = {}
    for m in re.finditer(r"(?:int64_t|int|const char\*)\s+(vlmclip_\w+)\s*\(([^;]*?)\)\s*;", body, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        decls[name] = n
    return decls


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__

    __graft_entry__.build()  # compiles for sm_100a if the in-tree .so is stale (nvcc cross-compiles without a GPU)
    from vlm_clip_b200 import _native

    return _native.load()


def test_header_declares_the_expected_surface():
    d = _declarations()
    assert len(d) >= 24
    for must in ("vlmclip_gemm_bf16", "vlmclip_attention_fwd", "vlmclip_adapter_fwd", "vlmclip_adapter_bwd",
                 "vlmclip_clip_loss", "vlmclip_class_head", "vlmclip_adamw_clip_step", "vlmclip_layernorm_bf16"):
        assert must in d


def test_every_declared_symbol_is_exported_and_bound(lib):
    from vlm_clip_b200 import _native

    decls = _declarations()
    for name, nargs in decls.items():
        assert hasattr(lib, name), f"{name} declared in vlmclip.h but not exported by the library"
        assert name in _native.PROTOTYPES, f"{name} has no ctypes prototype"
        assert len(_native.PROTOTYPES[name][1]) == nargs, f"{name}: header has {nargs} args, binding has " \
                                                            f"{len(_native.PROTOTYPES[name][1])}"
    assert set(_native.PROTOTYPES) == set(decls), set(_native.PROTOTYPES) ^ set(decls)


def test_abi_version_and_error_channel(lib):
    assert lib.vlmclip_abi_version() == 5
    assert isinstance(lib.vlmclip_last_error(), bytes)
    assert lib.vlmclip_launch_count() >= 0
    # argument validation happens before any CUDA call, so it can be exercised without a GPU
    rc = lib.vlmclip_gemm_bf16(None, 0, None, 0, None, 0, None, None, 0, None, None, None, 0, 1e-5, None, None, None, 1, 1, 1, 0, 0, None)
    assert rc < 0 and b"null" in lib.vlmclip_last_error()
    assert lib.vlmclip_adapter_bwd_workspace(5, 768, 256) == 8 * (3 * 768 + 2 * 256)
    assert lib.vlmclip_clip_loss_workspace(256, 512) == 4 * 256 + 256 * 256 + 2 * 256 * 512


def test_size_queries_and_argument_checks_without_gpu(lib):
    """Workspace / counter size queries are pure host arithmetic and the argument checks run before any CUDA call:
    both pin the dispatch rules (which sequence lengths take the key-range split) without a device."""
    import os

    if os.environ.get("VLMCLIP_ATTN_SPLIT") is None:  # auto: S <= 224 one tcgen05 launch, <= 288 mma.sync, <= 384 split
        assert lib.vlmclip_attention_fwd_workspace(4, 197, 12) == 0
        assert lib.vlmclip_attention_fwd_workspace(4, 257, 16) == 0
        assert lib.vlmclip_attention_fwd_workspace(4, 300, 16) == 2 * 4 * 300 * 16
        assert lib.vlmclip_attention_fwd_workspace(4, 385, 16) == 0
    assert lib.vlmclip_attention_fwd_workspace(0, 257, 16) == 0
    assert lib.vlmclip_clip_loss_counters(512) == 21 * 4 + 4  # 21 counters per 128-row block + 4
    rc = lib.vlmclip_attention_fwd(None, None, None, 2, 77, 8, 1, 0.125, None)
    assert rc < 0 and b"null" in lib.vlmclip_last_error()
    buf = ctypes.create_string_buffer(64)  # a non-null, 16-byte aligned host address: rejected by the S limit first
    addr = (ctypes.addressof(buf) + 15) & ~15
    rc = lib.vlmclip_attention_fwd(addr, addr, None, 2, 513, 8, 0, 0.125, None)
    assert rc < 0 and b"512" in lib.vlmclip_last_error()
    rc = lib.vlmclip_attention_fwd(addr + 2, addr, None, 2, 77, 8, 0, 0.125, None)
    assert rc < 0 and b"aligned" in lib.vlmclip_last_error()


def test_library_is_sm100a_tcgen05(lib):
    """The shipped cubin must contain the Blackwell tensor-core / TMA / TMEM instructions (cuobjdump mnemonics of
    B200_PROFILING.md), i.e. nothing silently fell back to a legacy path."""
    import shutil
    import subprocess

    from vlm_clip_b200 import _native

    cu = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cu).exists():
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cu, "-sass", str(_native.lib_path())], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM"):
        assert mnemonic in sass, mnemonic
