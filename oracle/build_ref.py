#!/usr/bin/env python
"""Build oracle/_ref/: the parts of the reference that compile in this container.

TEST INFRASTRUCTURE ONLY.  Compiles, from the sources where they lie under /root/reference
(never copied into the repo; outputs only under the git-ignored oracle/_ref/):

  libref_camera.so   src/vkgs/scene/camera.cc (+ vendored glm) behind oracle/ref_shim/camera_shim.cc
  libref_sort.so     third_party/vulkan_radix_sort/bench/cpu_benchmark.cc (the author's CPU statement of the
                     sort: std::stable_sort by key) behind oracle/ref_shim/sort_shim.cc
  libref_shaders.so  src/shader/{parse_ply,rank,inverse_index,projection}.comp, splat.vert, splat.frag executed
                     through glm by oracle/ref_shim/shader_harness.cc.  GLSL is turned into C++ by a purely
                     syntactic pass (below): preprocessor directives dropped, interface blocks flattened to
                     globals (unsized arrays -> pointers), swizzles spelled as glm calls, main -> shader_main.

The reference's own build (CMake + Vulkan SDK + glslangValidator + slangc) is not runnable here;
the Vulkan fixed-function stages and the Slang sort kernels therefore stay outside oracle/_ref.
"""
from __future__ import annotations

import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("VKGS_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
GEN = os.path.join(OUT, "gen")
SHIM = os.path.join(HERE, "ref_shim")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
CXXFLAGS = ["-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-fvisibility=hidden", "-w"]

SHADERS = ["parse_ply.comp", "rank.comp", "inverse_index.comp", "projection.comp", "splat.vert", "splat.frag"]


def glsl_to_cpp(src: str) -> str:
    src = re.sub(r"^\s*#\s*(version|extension|pragma)\b.*$", "", src, flags=re.M)
    src = re.sub(r"layout\s*\(\s*local_size[^)]*\)\s*in\s*;", "", src)

    def block(m):
        body = m.group("body")
        return re.sub(r"(\w+)\s+(\w+)\s*\[\s*\]\s*;", r"\1* \2;", body)

    src = re.sub(
        r"layout\s*\([^)]*\)\s*(?:(?:readonly|writeonly)\s+)*(?:uniform|buffer)\s+\w+\s*\{(?P<body>[^}]*)\}\s*;",
        block, src, flags=re.S)
    src = re.sub(r"layout\s*\(\s*location\s*=\s*\d+\s*\)\s*(?:in|out)\s+", "", src)
    src = re.sub(r"\bshared\s+", "", src)
    # swizzle store  X.xyz = e;  ->  component stores
    src = re.sub(r"^(\s*)(\S.*?)\.xyz\s*=\s*(.+?);\s*$",
                 r"\1{ vec3 _t = \3; \2.x = _t.x; \2.y = _t.y; \2.z = _t.z; }", src, flags=re.M)
    # swizzle loads -> glm's function swizzles
    src = re.sub(r"\.(xyz|xy|zw|yz|rgb)\b(?!\s*\()", r".\1()", src)
    src = re.sub(r"\bvoid\s+main\s*\(\s*\)", "void shader_main()", src)
    return src


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise SystemExit(1)


def build() -> bool:
    if not os.path.isdir(REF):
        print(f"[build_ref] {REF} not present: keeping prebuilt oracle/_ref as is")
        return False
    os.makedirs(GEN, exist_ok=True)
    glm = os.path.join(REF, "third_party", "glm")
    run([CXX, *CXXFLAGS, "-I", os.path.join(REF, "include"), "-I", glm,
         os.path.join(REF, "src", "vkgs", "scene", "camera.cc"), os.path.join(SHIM, "camera_shim.cc"),
         "-o", os.path.join(OUT, "libref_camera.so")])
    bench = os.path.join(REF, "third_party", "vulkan_radix_sort", "bench")
    # -include cstdint: benchmark_base.h uses uint32_t without including it (newer libstdc++ no longer leaks it)
    run([CXX, *CXXFLAGS, "-include", "cstdint", "-I", bench, os.path.join(bench, "cpu_benchmark.cc"), os.path.join(SHIM, "sort_shim.cc"),
         "-o", os.path.join(OUT, "libref_sort.so")])
    for s in SHADERS:
        with open(os.path.join(REF, "src", "shader", s)) as f:
            cpp = glsl_to_cpp(f.read())
        with open(os.path.join(GEN, s + ".inc"), "w") as f:
            f.write(f"// GENERATED from {REF}/src/shader/{s} by oracle/build_ref.py - do not commit\n" + cpp)
    run([CXX, *CXXFLAGS, "-I", glm, "-I", OUT, os.path.join(SHIM, "shader_harness.cc"),
         "-o", os.path.join(OUT, "libref_shaders.so")])
    print("[build_ref] built", ", ".join(sorted(f for f in os.listdir(OUT) if f.endswith(".so"))))
    return True


if __name__ == "__main__":
    build()
