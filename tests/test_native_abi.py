"""CPU: the C-ABI library builds, loads, and exports exactly what include/bjx.h declares (no compute calls)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

from bijax_b200 import _native as nv

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "bjx.h"


def _declared():
    txt = HEADER.read_text()
    return sorted(set(re.findall(r"BJX_API\s+[\w\s\*]+?\b(bjx_\w+)\s*\(", txt)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = nv.load()
    declared = _declared()
    assert declared, "no BJX_API declarations parsed"
    assert sorted(nv.EXPORTED_SYMBOLS) == declared
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.bjx_abi_version() == nv.ABI_VERSION


def test_struct_layouts_match_the_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "bjx.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(BjxElboArgs), sizeof(BjxCoord), offsetof(BjxElboArgs, key),'
        " offsetof(BjxElboArgs, workspace_bytes), offsetof(BjxElboArgs, ll_scale), offsetof(BjxElboArgs, kernel_override),"
        " offsetof(BjxElboArgs, x_planes), offsetof(BjxElboArgs, key_index), offsetof(BjxElboArgs, S_total),"
        " offsetof(BjxElboArgs, point_mode), offsetof(BjxElboArgs, rank), offsetof(BjxElboArgs, row_index), offsetof(BjxElboArgs, N_full));return 0;}\n"
    )
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [C.sizeof(nv.BjxElboArgs), C.sizeof(nv.BjxCoord), nv.BjxElboArgs.key.offset, nv.BjxElboArgs.workspace_bytes.offset,
            nv.BjxElboArgs.ll_scale.offset, nv.BjxElboArgs.kernel_override.offset, nv.BjxElboArgs.x_planes.offset,
            nv.BjxElboArgs.key_index.offset, nv.BjxElboArgs.S_total.offset, nv.BjxElboArgs.point_mode.offset, nv.BjxElboArgs.rank.offset,
            nv.BjxElboArgs.row_index.offset, nv.BjxElboArgs.N_full.offset]
    assert got == want
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "bjx.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu\\n", sizeof(BjxMinibatch), offsetof(BjxMinibatch, ldx_full), offsetof(BjxMinibatch, x_planes_full),'
        " offsetof(BjxMinibatch, planes_b), offsetof(BjxMinibatch, elbo_key));return 0;}\n"
    )
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    M = nv.BjxMinibatch
    assert got == [C.sizeof(M), M.ldx_full.offset, M.x_planes_full.offset, M.planes_b.offset, M.elbo_key.offset]


def test_header_is_plain_c():
    """The boundary must be consumable from C (cgo / JNI / ctypes style bindings): compile it as C99 with -pedantic."""
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", str(HEADER)], check=True)


def test_argument_validation_happens_before_any_cuda_work():
    lib = nv.load()
    a = nv.BjxElboArgs()
    a.S, a.N, a.D = 0, 10, 4  # S = 0 is invalid
    assert lib.bjx_elbo_workspace_bytes(C.byref(a)) == 0
    assert b"S" in lib.bjx_last_error()
    assert lib.bjx_elbo_step(None, None) == nv.ERR_INVALID_ARGUMENT
    with pytest.raises(nv.BjxError):
        nv.check(lib.bjx_random_normal(None, 4, 0, 4, 0, None, None))
    assert lib.bjx_minibatch_gather(None, None, None, 0, None) == nv.ERR_INVALID_ARGUMENT
    mb = nv.BjxMinibatch()
    mb.N_full, mb.B = 10, 4
    assert lib.bjx_minibatch_gather(C.byref(mb), None, None, 0, None) == nv.ERR_INVALID_ARGUMENT  # no key
    assert lib.bjx_train_steps_minibatch(None, None, 1, None, None, None, 0.1, 0.9, 0.999, 1e-8, None, None, None) == nv.ERR_INVALID_ARGUMENT
