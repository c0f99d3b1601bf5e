"""CPU: the C-ABI library builds for sm_100a, loads, and exports exactly what include/v5ela.h declares (no compute)."""
import ctypes
import os
import re
import subprocess

import pytest

from helpers import HERE

ROOT = os.path.dirname(HERE)
HEADER = os.path.join(ROOT, "include", "v5ela.h")


@pytest.fixture(scope="module")
def lib_path():
    from v5ela import build

    return build.build_library()


def declared_functions():
    with open(HEADER) as f:
        text = f.read()
    return sorted(set(re.findall(r"V5ELA_API[^;(]*?\b(v5ela_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    from v5ela import _abi

    assert declared_functions() == sorted(_abi.EXPORTS)


def test_library_exports_every_declared_symbol(lib_path):
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_path], text=True)
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l and "v5ela_" in l)
    assert exported == declared_functions()
    lib = ctypes.CDLL(lib_path)
    for name in declared_functions():
        assert hasattr(lib, name)


def test_library_carries_sm100a_sass_only(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, out


def test_host_only_entry_points(lib_path):
    from v5ela import _abi
    from v5ela.records import RECORD_BYTES

    lib = _abi.load()
    assert lib.v5ela_abi_version() == 1
    assert lib.v5ela_record_bytes() == RECORD_BYTES == 3144
    assert lib.v5ela_status_string(0) == b"ok"
    assert lib.v5ela_status_string(-3) == b"no sm_100 CUDA device"
    # C compilers agree on the record layout the header promises
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "v5ela.h"\nint main(void){printf("%zu %zu %zu %zu", sizeof(v5ela_record), ' \
          'offsetof(v5ela_record, ela_sum), offsetof(v5ela_record, tex_maxabs), offsetof(v5ela_record, ela_max));return 0;}'
    exe = os.path.join(ROOT, "tests", "emu", "_layout_check")
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src, text=True, check=True)
    assert subprocess.check_output([exe], text=True).split() == ["3144", "3072", "3136", "3138"]
    os.remove(exe)


def test_no_device_is_an_error_not_a_fallback(lib_path):
    import torch
    from v5ela import _abi

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_abi.V5ElaError) as ei:
        _abi.Handle(0)
    assert ei.value.status == -3
    import v5ela

    with pytest.raises(ValueError):
        v5ela.analyze_batch(torch.zeros((1, 16, 16, 3), dtype=torch.uint8))     # CPU tensor: refused, no CPU path


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fake-video-detection-engine_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    text = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), os.path.join(dirpath, fn)
                assert "libv5ela_oracle" not in text and "libv5ela_emu" not in text
