"""CPU: libgbcodec.so loads and exports every function include/gbcodec.h declares (no compute calls
— there is no GPU here), the ctypes structures have the C layout, and argument errors are reported
through the status/last_error convention."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gbcodec.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gbcodec_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from infantposeestimation_gaussianbias_b200 import _native
    return _native.lib()


def test_every_declared_symbol_is_exported_and_bound(lib):
    from infantposeestimation_gaussianbias_b200 import _native
    names = declared_functions()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/gbcodec.h but not exported by libgbcodec.so"
    assert set(names) == set(_native.EXPORTS), "the ctypes binding and the header disagree on the entry points"
    assert lib.gbcodec_abi_version() == _native.ABI_VERSION


def test_struct_layouts_match_the_header(tmp_path):
    """sizeof / offsetof from a C compile of the header against the ctypes mirrors."""
    from infantposeestimation_gaussianbias_b200 import _native as N
    prog = tmp_path / "layout.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "gbcodec.h"\n'
                    'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(gbcodec_loss_desc), offsetof(gbcodec_loss_desc, target_sigma),'
                    ' offsetof(gbcodec_loss_desc, pairs), sizeof(gbcodec_postprocess_desc), offsetof(gbcodec_postprocess_desc, input_h),'
                    ' sizeof(gbcodec_combined_desc), offsetof(gbcodec_combined_desc, heatmap_scale), offsetof(gbcodec_combined_desc, w_reg));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [C.sizeof(N.LossDesc), N.LossDesc.target_sigma.offset, N.LossDesc.pairs.offset,
            C.sizeof(N.PostprocessDesc), N.PostprocessDesc.input_h.offset,
            C.sizeof(N.CombinedDesc), N.CombinedDesc.heatmap_scale.offset, N.CombinedDesc.w_reg.offset]
    assert got == want


def test_argument_errors_use_the_status_convention(lib):
    """Checks that fail before any CUDA call: no device needed."""
    from infantposeestimation_gaussianbias_b200 import _native as N
    assert lib.gbcodec_encode_f32(None, None, None, None, 1, 1, 8, 8, 32.0, 32.0, 2.0, None) == -1      # NULL pointer
    assert b"NULL" in lib.gbcodec_last_error()
    assert lib.gbcodec_decode_argmax_f32(None, 1, 1, 8, 6, 0, None, None, None, None) == -2             # W % 4
    assert lib.gbcodec_status_string(-2) == b"bad shape"
    assert lib.gbcodec_combined_loss_f32(None, *([None] * 11), None, 0, None) == -1
    d = N.CombinedDesc(1, 1, 8, 8, 0, 0, 0, 1, 1, 1.0, 1.0, 0.5, 1.0, 0.1, 0.5)
    assert lib.gbcodec_combined_loss_f32(d, *([None] * 11), None, 0, None) == -4                        # terms = 0
    assert lib.gbcodec_loss_workspace_bytes(4, 17, 64, 48) >= 4 * 17 * 52
    assert lib.gbcodec_combined_workspace_bytes(4, 17) >= 4 * 17 * 16
    assert lib.gbcodec_combined_workspace_bytes(0, 17) == 0
