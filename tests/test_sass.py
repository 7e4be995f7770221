"""What the compiler made of the intersection loops (CPU test: reads the SASS of the built objects with cuobjdump).

DESIGN.md §4 quotes instruction counts for the always-executed path of a FILTER group (8 spheres) and of a triangle pair,
and states which instructions carry the work (FFMA2 / FMUL2 / FADD2, LDS, UBLKCP).  These properties were reached by
reading SASS; a compiler or source change that silently loses them (the re-derived shared-memory base address of
profiles/r02_bench.md, a contracted multiply-add in the exact kernel) should fail here, not be discovered on the GPU.
"""
import collections
import importlib.util
import shutil
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
OBJ = ROOT / "rust-swift-raytracer_b200" / "lib" / "obj"

spec = importlib.util.spec_from_file_location("sass_loops", ROOT / "scripts" / "sass_loops.py")
sass = importlib.util.module_from_spec(spec)
spec.loader.exec_module(sass)

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not installed")

# rt_render_kernel<FAST, SMEM, BLOCK, SPH, TRIS, NP> as mangled template arguments
C2 = "ILb0ELb1ELi256ELi0ELb0ELi1E"     # exact, staged, 256 threads, direct walk, no triangles   (8 spheres)
C3 = "ILb0ELb1ELi256ELi1ELb0ELi1E"     # exact, staged, 256 threads, FILTER walk, no triangles   (1,000 spheres)
C5 = "ILb0ELb1ELi1024ELi1ELb1ELi1E"    # exact, staged, 1024 threads, FILTER walk, triangles     (8,000 + 2,000)


@pytest.fixture(scope="module")
def exact():
    obj = OBJ / "rt_kernels_exact.o"
    if not obj.exists():
        import __graft_entry__
        __graft_entry__.build()
    return sass.disassemble(str(obj))


def kernel(funcs, args):
    names = [n for n in funcs if "rt_render_kernel" + args in n]
    assert len(names) == 1, names
    return funcs[names[0]]


def mix(instructions):
    return collections.Counter(sass.opcode(t) for _, t in instructions)


@pytest.mark.parametrize("args,limit", [(C3, 47), (C5, 49)])
def test_filter_group_always_path(exact, args, limit):
    """8 spheres against one ray: 8 LDS.128, 28 FFMA2, the max tree, one compare-and-branch, three loop instructions."""
    ins = kernel(exact, args)
    loop = sass.find_loop(ins, {"FFMA2": 28, "LDS": 16})       # 8 loads of the filter + 8 of the survivors' centres
    assert loop is not None
    path = sass.hot_path(ins, loop[0])
    c = mix(path)
    assert c["FFMA2"] == 28 and c["LDS"] == 8
    assert c["S2UR"] == 0 and c["S2R"] == 0, "the staged block's address is re-derived inside the loop"
    assert c["FFMA"] == 0 and c["FMUL"] == 0 and c["FADD"] == 0, "scalar arithmetic on the always-executed path"
    assert len(path) <= limit, f"{len(path)} instructions per FILTER group (DESIGN.md: {limit})"


def test_triangle_pair_always_path(exact):
    ins = kernel(exact, C5)
    loop = sass.find_loop(ins, {"FFMA2": 13, "FMUL2": 7})
    assert loop is not None
    path = sass.hot_path(ins, loop[0])
    c = mix(path)
    assert c["S2UR"] == 0 and c["S2R"] == 0
    assert c["MUFU"] == 2 and c["LDS"] == 2
    assert len(path) <= 34, f"{len(path)} instructions per triangle pair (DESIGN.md: 34)"


def test_exact_kernels_keep_the_two_roundings(exact):
    """--fmad=false: no scalar multiply-add may appear in the discriminants; the two-wide sums go through f2_add1
    (FFMA2 with the opaque 1.0), so the C2 kernel must still hold packed multiplies AND packed adds."""
    c = mix(kernel(exact, C2))
    assert c["FMUL2"] >= 28 and c["FADD2"] >= 16 and c["FFMA2"] >= 20
    for args in (C2, C3, C5):
        ins = kernel(exact, args)
        assert mix(ins)["UBLKCP"] >= 1, "scene staging is not a TMA bulk copy"
        assert not any("WGMMA" in t or "HMMA" in t for _, t in ins)


def test_register_budgets(exact):
    """64 registers for 4 CTAs x 256 threads (and 1 x 1024), 80 for the FILTER kernels' 3 x 256."""
    log = (ROOT / "rust-swift-raytracer_b200" / "lib" / "build.log")
    if not log.exists():
        pytest.skip("no build log (library was not rebuilt in this checkout)")
    text = log.read_text()
    import re
    regs = {}
    for m in re.finditer(r"Function properties for (\S+)\s*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores.*\n.*Used (\d+) registers", text):
        regs[m.group(1)] = (int(m.group(4)), int(m.group(3)))
    def of(args):
        k = [n for n in regs if "rt_render_kernel" + args in n]
        if not k:
            pytest.skip("the last build did not recompile the kernels (no ptxas statistics in the log)")
        return regs[k[0]]
    assert of(C2)[0] <= 64 and of(C2)[1] == 0, of(C2)
    assert of(C3)[0] <= 80 and of(C3)[1] <= 16, of(C3)
    assert of(C5)[0] <= 64, of(C5)
