"""Source rule behind programmatic dependent launch (csrc/common.cuh: launch_k): a kernel may be launched with the
programmatic-stream-serialization attribute only if its first statement is pdl_enter() (griddepcontrol.wait before any
global access), otherwise it could read what the previous grid has not written yet. Checked on the sources, no GPU."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gi-gs_b200", "csrc")


def _sources():
    return "\n".join(open(f).read() for f in sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh"))))


def test_every_kernel_launched_with_launch_k_waits_first():
    txt = _sources()
    names = {m.group(1) for m in re.finditer(r"launch_k\(\s*\(?([A-Za-z_][A-Za-z_0-9]*)", txt)}
    names -= {"void", "kern"}                 # the helper's own definition; gi_march.cu's function-pointer variable
    names.add("gi_march_kernel")              # ... which is always an instance of this template
    assert len(names) >= 40
    for n in sorted(names):
        defs = list(re.finditer(r"__global__[^;{]*?\b" + n + r"\s*\([^;{]*?\)\s*\{\s*(\S[^\n]*)", txt, re.S))
        assert defs, f"{n}: launched with launch_k but no __global__ definition found"
        for d in defs:
            assert d.group(1).strip().startswith("pdl_enter();"), f"{n}: first statement is not pdl_enter()"


def test_no_kernel_with_pdl_enter_is_left_without_a_reason():
    """the other direction, informational: kernels that wait but are launched the ordinary way are fine (the wait is a
    no-op then); the count documents how many kernels chain"""
    txt = _sources()
    assert txt.count("pdl_enter();") >= 40
