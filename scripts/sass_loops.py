#!/usr/bin/env python3
"""SASS view of the render kernels' inner loops (no GPU needed: reads cuobjdump's disassembly).

    python scripts/sass_loops.py [object-or-library] [--kernel SUBSTR] [--hot HEAD_HEX]

Lists every backward branch (loop) of each rt_render_kernel instantiation with its length and opcode mix; with --hot,
prints the loop's always-executed path: from the loop head, every predicated forward branch is taken (that is the path of a
group of primitives none of which survives its filter), up to the backward branch.  tests/test_sass.py keeps the two numbers
DESIGN.md quotes (instructions per FILTER group and per triangle pair) from regressing.
"""
from __future__ import annotations

import collections
import re
import subprocess
import sys
from pathlib import Path

INS = re.compile(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*;?\s*/\*")
BRA = re.compile(r"^(@!?U?P\d\s+)?BRA(?:\.\w+)*\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)")


def disassemble(path: str) -> dict[str, list[tuple[int, str]]]:
    """{mangled kernel name: [(address, instruction text)]} of every function in the object / library."""
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    funcs: dict[str, list[tuple[int, str]]] = {}
    cur = None
    for line in out.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
            funcs[cur] = []
            continue
        m = INS.match(line)
        if m and cur is not None:
            funcs[cur].append((int(m.group(1), 16), m.group(2).strip().rstrip(";").strip()))
    return funcs


def opcode(text: str) -> str:
    parts = text.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    return op.split(".")[0]


def loops(ins: list[tuple[int, str]]):
    """[(head address, tail address, body)] for every backward branch."""
    index = {a: i for i, (a, _) in enumerate(ins)}
    found = []
    for i, (a, t) in enumerate(ins):
        m = BRA.match(t)
        if m:
            tgt = int(m.group(2), 16)
            if tgt <= a and tgt in index:
                found.append((tgt, a, ins[index[tgt]:i + 1]))
    return found


def hot_path(ins: list[tuple[int, str]], head: int) -> list[tuple[int, str]]:
    """The path from `head` that takes every predicated forward branch, up to and including the first backward branch."""
    index = {a: i for i, (a, _) in enumerate(ins)}
    i, path = index[head], []
    while True:
        a, t = ins[i]
        path.append((a, t))
        m = BRA.match(t)
        if m:
            tgt = int(m.group(2), 16)
            if tgt <= a:
                return path
            i = index[tgt]
            continue
        i += 1


def find_loop(ins, want: dict[str, int]):
    """The innermost loop whose body holds exactly want[op] instructions of each given opcode (e.g. {'FFMA2': 28})."""
    best = None
    for head, tail, body in loops(ins):
        c = collections.Counter(opcode(t) for _, t in body)
        if all(c[k] == v for k, v in want.items()) and (best is None or len(body) < len(best[2])):
            best = (head, tail, body)
    return best


def main(argv):
    lib = str(Path(__file__).resolve().parent.parent / "rust-swift-raytracer_b200" / "lib" / "obj" / "rt_kernels_exact.o")
    kernel, hot = "rt_render_kernel", None
    args = list(argv)
    while args:
        a = args.pop(0)
        if a == "--kernel": kernel = args.pop(0)
        elif a == "--hot": hot = int(args.pop(0), 16)
        else: lib = a
    for name, ins in disassemble(lib).items():
        if kernel not in name or not ins:
            continue
        print(f"== {name}: {len(ins)} instructions")
        if hot is not None:
            path = hot_path(ins, hot)
            for a, t in path:
                print(f"   {a:05x} {t}")
            print(f"   always-executed path: {len(path)} instructions")
            continue
        for head, tail, body in loops(ins):
            c = collections.Counter(opcode(t) for _, t in body)
            print(f"   loop {head:#x}..{tail:#x} len {len(body):4d}  " + " ".join(f"{k}:{v}" for k, v in c.most_common(10)))


if __name__ == "__main__":
    main(sys.argv[1:])
