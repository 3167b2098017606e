"""TEST INFRASTRUCTURE ONLY.  Builds tests/emu/_build/libband_emu.so: the kernel part of
cmda_b200/csrc/voxel_factored.cu (everything above its host launch section), compiled for the HOST against the
fiber-based stand-in for CUDA in tests/emu/include/cuda_runtime.h.  The sources are copied with three mechanical
edits: `extern __shared__` -> `extern` (the harness defines the arrays), the three inline-PTX load / store helpers of
common.cuh -> plain loads / stores, and the header include path."""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "cmda_b200", "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libband_emu.so")
CUT = "// ---- workspace + launch sequence"


def _transform_common(src: str) -> str:
    src = src.replace('#include "../../include/cmda_b200.h"', f'#include "{os.path.join(ROOT, "include", "cmda_b200.h")}"')
    # inline PTX -> plain memory operations: the three helpers are replaced whole
    def swap(name, body):
        nonlocal src
        pat = re.compile(r"(__device__ __forceinline__ [^\n]*\b" + name + r"\([^)]*\) \{).*?\n\}\n", re.S)
        assert pat.search(src), name
        src = pat.sub(lambda m: m.group(1) + "\n    " + body + "\n}\n", src, count=1)
    swap("ldg_stream_u4", "return *static_cast<const uint4*>(p);")
    swap("ldg_stream_u2", "return *static_cast<const uint2*>(p);")
    swap("stg_stream_f4", "*reinterpret_cast<float4*>(p) = v;")
    assert "asm" not in src, "an inline-PTX helper of common.cuh is not covered by the emulation build"
    return src


def generate() -> None:
    gen = os.path.join(BUILD, "gen")
    os.makedirs(gen, exist_ok=True)
    with open(os.path.join(CSRC, "common.cuh")) as f:
        common = _transform_common(f.read())
    with open(os.path.join(gen, "common.cuh"), "w") as f:
        f.write(common)
    with open(os.path.join(CSRC, "event_math.cuh")) as f:
        em = f.read()
    with open(os.path.join(gen, "event_math.cuh"), "w") as f:
        f.write(em)
    with open(os.path.join(CSRC, "voxel_factored.cu")) as f:
        vf = f.read()
    assert CUT in vf
    vf = vf[: vf.index(CUT)] + "\n}  // namespace cmda\n"
    vf = re.sub(r"extern __shared__( __align__\(\d+\))?", "extern", vf)
    assert "asm" not in vf.replace("masm", "")
    with open(os.path.join(gen, "voxel_factored_kernels.inc"), "w") as f:
        f.write(vf)


def build(force: bool = False) -> str:
    srcs = [os.path.join(CSRC, n) for n in ("common.cuh", "event_math.cuh", "voxel_factored.cu")] + \
           [os.path.join(HERE, "band_emu.cpp"), os.path.join(HERE, "include", "cuda_runtime.h"), os.path.abspath(__file__)]
    if not force and os.path.isfile(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return LIB
    generate()
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-w",
           "-I", os.path.join(HERE, "include"), "-I", BUILD, "-o", LIB, os.path.join(HERE, "band_emu.cpp")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("emulation build failed:\n" + res.stderr[-6000:])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
