"""TEST INFRASTRUCTURE ONLY.  Builds tests/emu/_build/libband_emu.so: the kernel part of
cmda_b200/csrc/voxel_factored.cu (everything above its host launch section), compiled for the HOST against the
fiber-based stand-in for CUDA in tests/emu/include/cuda_runtime.h.  The sources are copied with three mechanical
edits: `extern __shared__` -> `extern` (the harness defines the arrays), the three inline-PTX load / store helpers of
common.cuh -> plain loads / stores, and the header include path."""
import os
import re
import shlex
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "cmda_b200", "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libband_emu.so")
CUT = "// ---- workspace + launch sequence"
EXTRA = shlex.split(os.environ.get("EMU_EXTRA_FLAGS", ""))      # e.g. -fsanitize=address (tools/emu_asan.sh)


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
    swap("dependency_wait", ";")            # programmatic dependent launch: the emulation runs kernels one after another
    swap("dependency_release", ";")
    # ... and launch_dependent() is an ordinary launch of the stand-in
    pat = re.compile(r"(inline cudaError_t launch_dependent\(.*?\) \{)\n.*?\n\}\n", re.S)
    assert pat.search(src), "launch_dependent"
    src = pat.sub(lambda m: m.group(1) + "\n    (void)shm; (void)st;\n    emu_launch(grid, block, [&] { kernel(static_cast<KArgs>(args)...); });"
                  "\n    return cudaSuccess;\n}\n", src, count=1)
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


def _split_top(text: str):
    """Split at top-level commas (brackets of any kind nest)."""
    parts, depth, cur = [], 0, ""
    for ch in text:
        if ch in "([{<":
            depth += 1
        elif ch in ")]}>":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    return parts


def transform_launches(src: str) -> str:
    """`kernel<T...><<<grid, block, shm, stream>>>(args);` -> `emu_launch(grid, block, [&] { kernel<T...>(args); });`"""
    out, pos = "", 0
    pat = re.compile(r"([A-Za-z_][A-Za-z0-9_]*(?:<[^<>;(){}]*>)?)<<<")
    while True:
        m = pat.search(src, pos)
        if not m:
            return out + src[pos:]
        close = src.index(">>>", m.end())
        cfg = _split_top(src[m.end():close])
        assert len(cfg) in (2, 3, 4), cfg
        # the argument list: from the '(' after >>> to its matching ')'
        a0 = close + 3
        assert src[a0] == "(", src[a0:a0 + 20]
        depth, i = 0, a0
        while True:
            if src[i] == "(":
                depth += 1
            elif src[i] == ")":
                depth -= 1
                if depth == 0:
                    break
            i += 1
        args = src[a0 + 1:i]
        out += src[pos:m.start()] + f"emu_launch({cfg[0].strip()}, {cfg[1].strip()}, [&] {{ {m.group(1)}({args}); }})"
        pos = i + 1


def generate_full() -> None:
    """Whole translation units (kernels AND their host launch code) for the emulated voxel path."""
    generate()
    gen = os.path.join(BUILD, "gen")
    for name in ("voxel_factored.cu", "norm.cu"):
        with open(os.path.join(CSRC, name)) as f:
            src = f.read()
        src = re.sub(r"extern __shared__( __align__\(\d+\))?", "extern", src)
        src = transform_launches(src)
        assert "<<<" not in src and "asm" not in src.replace("masm", "")
        with open(os.path.join(gen, name.replace(".cu", ".cpp")), "w") as f:
            f.write(src)


def build_vg(force: bool = False) -> str:
    """tests/emu/_build/libvg_emu.so: launch_factored + launch_norm_apply (the real host launch code and every
    kernel under it) on the emulation, behind emu_events_vg (vg_emu.cpp)."""
    lib = os.path.join(BUILD, "libvg_emu.so")
    srcs = [os.path.join(CSRC, n) for n in ("common.cuh", "event_math.cuh", "voxel_factored.cu", "norm.cu")] + \
           [os.path.join(HERE, "vg_emu.cpp"), os.path.join(HERE, "include", "cuda_runtime.h"), os.path.abspath(__file__)]
    if not force and os.path.isfile(lib) and all(os.path.getmtime(lib) >= os.path.getmtime(s) for s in srcs):
        return lib
    generate_full()
    gen = os.path.join(BUILD, "gen")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-w"] + EXTRA + \
          ["-I", os.path.join(HERE, "include"), "-I", gen] + \
          ["-o", lib, os.path.join(HERE, "vg_emu.cpp"), os.path.join(gen, "voxel_factored.cpp"), os.path.join(gen, "norm.cpp")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("emulation build failed:\n" + res.stderr[-8000:])
    return lib


def _transform_tiled(src: str) -> str:
    """voxel_tiled.cu: `atom.shared.add.u32` / predicated `red.shared.add.s32` on 32-bit shared-window addresses ->
    the same read-modify-writes through emu_shared_window (abi_emu.cpp points it at the kernel's shared array)."""
    a0 = 'if constexpr (XOFF == 0) asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(q) : "memory");'
    a1 = 'else asm volatile("atom.shared.add.u32 %0, [%1+4], %2;" : "=r"(old) : "r"(addr), "r"(q) : "memory");'
    assert a0 in src and a1 in src
    src = src.replace(a0, "old = atomicAdd(reinterpret_cast<unsigned*>(emu_shared_window + addr + XOFF), static_cast<unsigned>(q));")
    src = src.replace(a1, "")
    r0 = src.index("    if constexpr (XOFF == 0)\n        asm volatile(\"{ .reg .pred p;")
    r1 = src.index("\n}", r0)
    src = src[:r0] + "    if (d != 0) atomicAdd(reinterpret_cast<int*>(emu_shared_window + haddr + XOFF), d);" + src[r1:]
    return src


FULL_UNITS = ("api.cu", "slicer.cu", "voxel_global.cu", "voxel_tiled.cu", "voxel_factored.cu", "voxel_exact.cu", "norm.cu",
              "pseudo_events.cu", "resize.cu")


def build_abi(force: bool = False) -> str:
    """tests/emu/_build/libcmda_b200_emu.so: the C ABI of include/cmda_b200.h with every translation unit compiled
    against the emulation.  Two more mechanical edits make that possible: TILED's two inline-PTX shared-memory atomics
    become the C++ they abbreviate (shared-window addresses are offsets into the kernel's shared array), and EXACT's
    cub::DeviceRadixSort::SortPairs resolves to a stable host sort (include/cub/).  "Device" pointers are host pointers."""
    lib = os.path.join(BUILD, "libcmda_b200_emu.so")
    srcs = [os.path.join(CSRC, n) for n in ("common.cuh", "event_math.cuh") + FULL_UNITS] + \
           [os.path.join(HERE, "abi_emu.cpp"), os.path.join(HERE, "include", "cuda_runtime.h"), os.path.abspath(__file__),
            os.path.join(ROOT, "include", "cmda_b200.h")]
    if not force and os.path.isfile(lib) and all(os.path.getmtime(lib) >= os.path.getmtime(s) for s in srcs):
        return lib
    generate()
    gen = os.path.join(BUILD, "gen")
    objs, procs = [], []
    flags = ["-std=c++17", "-O1", "-g", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-w"] + EXTRA + \
            ["-I", os.path.join(HERE, "include"), "-I", gen]
    for name in FULL_UNITS:
        with open(os.path.join(CSRC, name)) as f:
            src = f.read()
        src = re.sub(r"extern __shared__( __align__\(\d+\))?", "extern", src)
        src = transform_launches(src)
        if name == "voxel_tiled.cu":
            src = _transform_tiled(src)
        assert "<<<" not in src and "asm" not in src.replace("masm", ""), name
        cpp = os.path.join(gen, name.replace(".cu", ".cpp"))
        with open(cpp, "w") as f:
            f.write(src)
        obj = os.path.join(gen, name.replace(".cu", ".o"))
        objs.append(obj)
        procs.append((name, subprocess.Popen(["g++"] + flags + ["-c", "-o", obj, cpp], stderr=subprocess.PIPE, text=True)))
    for name, pr in procs:
        err = pr.communicate()[1]
        if pr.returncode != 0:
            raise RuntimeError(f"emulation build of {name} failed:\n" + err[-8000:])
    res = subprocess.run(["g++"] + flags + ["-shared", "-o", lib, os.path.join(HERE, "abi_emu.cpp")] + objs, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("emulation link failed:\n" + res.stderr[-8000:])
    return lib


def build(force: bool = False) -> str:
    srcs = [os.path.join(CSRC, n) for n in ("common.cuh", "event_math.cuh", "voxel_factored.cu")] + \
           [os.path.join(HERE, "band_emu.cpp"), os.path.join(HERE, "include", "cuda_runtime.h"), os.path.abspath(__file__)]
    if not force and os.path.isfile(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return LIB
    generate()
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-w"] + EXTRA + \
          ["-I", os.path.join(HERE, "include"), "-I", BUILD, "-o", LIB, os.path.join(HERE, "band_emu.cpp")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("emulation build failed:\n" + res.stderr[-6000:])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
    print(build_vg(force=True))
    print(build_abi(force=True))
