"""glsl_front.py — build-time syntax adapter: reference GLSL (read where it lies) -> C++-parsable text in oracle/_ref/gen/.

TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).  Used by oracle/ref_shim/Makefile to build oracle/_ref/libigxref_*.so, the
reference's OWN shader code compiled for the host, against which the hand-written oracle is pinned.

Nothing of the reference is stored in this repository: the inputs are read from <reference>/res/shaders at build time and the
outputs go to oracle/_ref/gen/ (git-ignored).  Every arithmetic statement, branch, loop and function body passes through
byte for byte; only constructs that C++ has no spelling for are rewritten, line by line, keeping line numbers:

  1. `#version`, `#extension` lines and `layout(local_size_* ...) in;` are blanked (no C++ meaning; the harness iterates the grid);
  2. interface blocks  `layout(...) [readonly|writeonly] uniform|buffer Name { T a; U b[]; };`  become the globals the block
     declares:  `static T a; static U* b;`  (bound by the harness to the host arrays);
  3. opaque uniforms   `layout(...) [writeonly] uniform sampler2D|sampler2DMS|image2D name;`  become  `static <type> name;`;
  4. `inout T x` parameters become `T& x` (GLSL copy-in/copy-out == a C++ reference for the non-aliased calls of this code);
  5. `.rrr` applied to a SCALAR expression (five uses, all inside composite.comp's non-default DEBUG views) becomes `*vec3(1)`.

The script prints how many lines each rule touched and fails if a file still contains a `layout(` it did not understand.
GLSL float literals (no suffix = 32-bit) are kept as written; the Makefile compiles with -fsingle-precision-constant.
"""
from __future__ import annotations

import os
import re
import sys

FILES = ["defines.glsl", "utils.glsl", "rand_util.glsl", "primitive.glsl", "camera.glsl", "scene.glsl", "trace.glsl", "light.glsl",
         "light_rt.glsl", "init.comp", "raygen.comp", "nv_all.shadow.comp", "nv_all.lighting.comp", "composite.comp"]

BLOCK = re.compile(r"layout\s*\(([^)]*)\)\s*((?:readonly|writeonly)\s+)?(uniform|buffer)\s+(\w+)\s*\{([^}]*)\}\s*;")
OPAQUE = re.compile(r"layout\s*\(([^)]*)\)\s*((?:readonly|writeonly)\s+)?uniform\s+(sampler2DMS|sampler2D|image2D)\s+(\w+)\s*;")
LOCAL_SIZE = re.compile(r"^[ \t]*layout\s*\(\s*local_size_[^)]*\)\s*in\s*;[ \t]*$", re.M)
DIRECTIVE = re.compile(r"^[ \t]*#[ \t]*(version|extension)\b.*$", re.M)
INOUT = re.compile(r"\binout\s+(\w+)\s+(\w+)")
MEMBER = re.compile(r"\s*([\w]+)\s+(\w+)\s*(\[\s*\])?\s*;")
SCALAR_RRR = re.compile(r"(\)|\btransparency)\.rrr\b")


def keep_lines(old: str, new: str) -> str:
    """pad `new` with the newlines `old` had so that line numbers stay those of the reference file"""
    return new + "\n" * max(0, old.count("\n") - new.count("\n"))


def adapt(text: str, stats: dict) -> str:
    def block(m):
        stats["interface blocks"] += 1
        decls = []
        body = re.sub(r"//[^\n]*", "", m.group(5))
        for mem in MEMBER.finditer(body):
            ty, name, arr = mem.group(1), mem.group(2), mem.group(3)
            decls.append(f"static {ty}* {name};" if arr else f"static {ty} {name};")
        return keep_lines(m.group(0), " ".join(decls))

    def opaque(m):
        stats["opaque uniforms"] += 1
        return f"static {m.group(3)} {m.group(4)};"

    text, n = DIRECTIVE.subn("", text)
    stats["#version/#extension"] += n
    text, n = LOCAL_SIZE.subn("", text)
    stats["local_size"] += n
    text = BLOCK.sub(block, text)
    text = OPAQUE.sub(opaque, text)
    text, n = INOUT.subn(r"\1& \2", text)
    stats["inout parameters"] += n
    text, n = SCALAR_RRR.subn(r"\1*vec3(1)", text)
    stats["scalar .rrr"] += n
    if re.search(r"\blayout\s*\(", text):
        raise SystemExit("glsl_front: a layout(...) declaration was not understood")
    return text


def main(src_dir: str, out_dir: str) -> None:
    os.makedirs(out_dir, exist_ok=True)
    stats = {k: 0 for k in ("#version/#extension", "local_size", "interface blocks", "opaque uniforms", "inout parameters", "scalar .rrr")}
    changed_total = 0
    for name in FILES:
        src = open(os.path.join(src_dir, name), encoding="utf-8", errors="replace").read()
        out = adapt(src, stats)
        a, b = src.split("\n"), out.split("\n")
        assert len(a) == len(b), f"{name}: line count changed"
        changed = sum(1 for x, y in zip(a, b) if x != y)
        changed_total += changed
        with open(os.path.join(out_dir, name), "w", encoding="utf-8") as f:
            f.write(f'#line 1 "{os.path.join(src_dir, name)}"\n' + out)
        print(f"  {name}: {len(a)} lines, {changed} adapted")
    print("glsl_front:", ", ".join(f"{v} {k}" for k, v in stats.items()), f"({changed_total} lines adapted in total)")


if __name__ == "__main__":
    if len(sys.argv) != 3:
        raise SystemExit("usage: glsl_front.py <reference>/res/shaders <out dir>")
    main(sys.argv[1], sys.argv[2])
