"""CPU suite: facts about the compiled sm_100a code of the streaming kernels that the design relies on, read from the
SASS with cuobjdump (no GPU needed): no local-memory spills, 128-bit accesses on the 16-byte paths, and no 64-bit division
subroutine inside the per-word loops (a `while` correction once compiled into one: DESIGN.md 4.8)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSTR = re.compile(r"^\s+/\*[0-9a-f]{4}\*/\s+(\S.*?);")


def _sass(obj):
    from lol_b200 import build
    build.build_library()
    path = os.path.join(build.OBJ_DIR, obj)
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    demangled = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
    funcs, cur, names = {}, None, iter(demangled)
    for line in out.splitlines():
        if "Function :" in line:
            cur = next(names)
            funcs[cur] = []
        else:
            m = INSTR.match(line)
            if m and cur is not None:
                funcs[cur].append(m.group(1))
    return funcs


pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or shutil.which("c++filt") is None, reason="needs cuobjdump and c++filt")


def _find(funcs, *needles):
    hits = [k for k in funcs if all(n in k for n in needles)]
    assert len(hits) == 1, (needles, list(funcs))
    return funcs[hits[0]]


def test_no_local_memory_in_streaming_kernels():
    for obj in ("ext_stream.o", "coeff_stream.o", "she_stream.o", "fused_stream.o"):
        for name, ins in _sass(obj).items():
            assert not any(re.match(r"(LDL|STL)\b", i) for i in ins), (obj, name)


def test_sixteen_byte_paths_use_128_bit_accesses():
    funcs = _sass("ext_stream.o")
    for word in ("WordZq2", "WordI64x2", "WordC64"):
        ins = _find(funcs, "k_ext_gather", word + ">")
        assert sum(".128" in i for i in ins if i.startswith("LDG")) >= 8, word
        assert all(".128" in i for i in ins if i.startswith("STG")), word


def test_coefficient_loops_have_no_division_call():
    """Two calls remain per kernel: the once-per-thread split of the global index into (coefficient, limb)."""
    funcs = _sass("coeff_stream.o")
    for op in ("OpLift", "OpReduce", "OpRescaleDropFast", "OpRescaleMod"):
        ins = _find(funcs, "k_coeff_stream", op + ">")
        assert sum(i.startswith("CALL") for i in ins) <= 2, op
