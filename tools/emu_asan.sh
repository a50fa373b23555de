#!/bin/bash
# AddressSanitizer pass over the kernel logic: the CPU tier's emulator tests against `make emu-asan` (exact-size device
# allocations, reused buffers poisoned past their logical size).  Test infrastructure; no GPU.
#   tools/emu_asan.sh [pytest args ...]      default: every test file that drives the emulator through ctypes
set -e
cd "$(dirname "$0")/.."
make -s emu-asan CXX=/usr/bin/g++
ASAN_RT=$(/usr/bin/g++ -print-file-name=libasan.so)
export B200SA_EMU_ASAN=1
export ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:abort_on_error=1:${ASAN_OPTIONS}
if [ $# -eq 0 ]; then
    set -- tests/test_emu_kernels.py tests/test_fuzz_emu.py tests/test_batch.py tests/test_lcp.py tests/test_group.py \
           tests/test_sharded_cpu.py tests/test_untrusted_bwt.py tests/test_wide.py
fi
LD_PRELOAD="$ASAN_RT" python -m pytest -x -q -m "not gpu" -p no:cacheprovider "$@"
