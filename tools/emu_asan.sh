#!/bin/bash
# The CPU emulation tests (tests/test_emu_*.py) with the emulated libraries rebuilt under AddressSanitizer: out-of-bounds
# reads / writes of the "device" buffers (exact-size numpy arrays and vectors) abort the run.  CPU only, about two minutes.
set -e
cd "$(dirname "$0")/.."
export EMU_EXTRA_FLAGS="-fsanitize=address -fno-omit-frame-pointer"
rm -rf tests/emu/_build
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
    python -m pytest tests/test_emu_banded.py tests/test_emu_voxel_path.py tests/test_emu_abi.py -x -q -p no:cacheprovider
rm -rf tests/emu/_build      # the next ordinary run rebuilds without the sanitizer
