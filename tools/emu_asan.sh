#!/bin/bash
# Sanitizer passes over the kernel logic: the CPU tier's emulator tests against a sanitizer build of the emulator library.
# Test infrastructure; no GPU.
#   tools/emu_asan.sh [asan|ubsan] [pytest args ...]
# asan  (default): exact-size device allocations, reused buffers poisoned past their logical size (make emu-asan)
# ubsan: shifts >= operand width, signed overflow, misaligned 64- / 128-bit vector accesses (make emu-ubsan)
set -e
cd "$(dirname "$0")/.."
VARIANT=asan
if [ "$1" = asan ] || [ "$1" = ubsan ]; then VARIANT=$1; shift; fi
make -s emu-$VARIANT CXX=/usr/bin/g++
export B200SA_EMU_SANITIZER=$VARIANT
if [ $# -eq 0 ]; then
    set -- tests/test_emu_kernels.py tests/test_fuzz_emu.py tests/test_batch.py tests/test_lcp.py tests/test_group.py \
           tests/test_sharded_cpu.py tests/test_untrusted_bwt.py tests/test_wide.py
fi
if [ $VARIANT = asan ]; then
    export ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:abort_on_error=1:${ASAN_OPTIONS}
    export LD_PRELOAD="$(/usr/bin/g++ -print-file-name=libasan.so)"
else
    export UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1:${UBSAN_OPTIONS}
    export LD_PRELOAD="$(/usr/bin/g++ -print-file-name=libubsan.so)"
fi
python -m pytest -x -q -m "not gpu" -p no:cacheprovider "$@"
