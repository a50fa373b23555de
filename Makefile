# Top-level build.  `python -c "import __graft_entry__ as g; g.build()"` runs `make all`.
#
#   lib      msufsort_b200/lib/libb200sa.so          hand-written CUDA kernels + C ABI (sm_100a only)
#   textgen  msufsort_b200/lib/libb200sa_textgen.so  synthetic input generators (host C)
#   facade   msufsort_b200/lib/libmsufsort.so        the reference-shaped C++ facade (src/library)
#   cli      msufsort_b200/lib/msufsort              command line tool with the reference demo's modes (src/executable)
#   facade_bench msufsort_b200/lib/facade_bench      end-to-end timing of the drop-in C++ path (tools/facade_bench.cpp; bench.py runs it)
#   oracle   oracle/liboracle.so (+ oracle/_ref/ when /root/reference exists)   TEST INFRASTRUCTURE
#   emu      tests/emu/libb200sa_emu.so              kernel logic under a CPU SIMT emulator (tests only)
NVCC      ?= nvcc
CC        ?= gcc
CXX       ?= g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCC_DEFS ?=
NVCCFLAGS := $(ARCH) $(NVCC_DEFS) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v --expt-relaxed-constexpr
CSRC      := msufsort_b200/csrc
LIBDIR    := msufsort_b200/lib
KSRC      := $(CSRC)/b200sa.cu $(CSRC)/comm.cuh $(CSRC)/engine_shard.inl $(CSRC)/engine_peer.inl $(CSRC)/engine_lcp.inl $(CSRC)/engine_batch.inl $(CSRC)/c_abi.inl $(CSRC)/c_abi_group.inl $(CSRC)/engine.cuh $(CSRC)/common.cuh $(CSRC)/radix_sort.cuh $(CSRC)/sa_kernels.cuh $(CSRC)/bwt_kernels.cuh $(CSRC)/lcp_kernels.cuh $(CSRC)/batch_kernels.cuh include/b200sa.h

all: lib textgen facade cli facade_bench oracle emu

lib: $(LIBDIR)/libb200sa.so
$(LIBDIR)/libb200sa.so: $(KSRC)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(CSRC)/b200sa.cu 2> $(LIBDIR)/ptxas.log || (cat $(LIBDIR)/ptxas.log; exit 1)
	@grep -E "error|warning" $(LIBDIR)/ptxas.log | grep -v "ptxas info" || true

textgen: $(LIBDIR)/libb200sa_textgen.so
$(LIBDIR)/libb200sa_textgen.so: $(CSRC)/textgen.c
	@mkdir -p $(LIBDIR)
	$(CC) -O2 -std=c11 -fPIC -fvisibility=hidden -shared -o $@ $<

facade: $(LIBDIR)/libmsufsort.so
$(LIBDIR)/libmsufsort.so: src/library/msufsort/msufsort.cpp src/library/msufsort/msufsort.h src/library/msufsort.h include/b200sa.h $(LIBDIR)/libb200sa.so
	$(CXX) -O2 -std=c++17 -fPIC -shared -Isrc -Iinclude -o $@ src/library/msufsort/msufsort.cpp -L$(LIBDIR) -lb200sa -Wl,-rpath,'$$ORIGIN'

cli: $(LIBDIR)/msufsort
$(LIBDIR)/msufsort: src/executable/msufsort/main.cpp $(LIBDIR)/libmsufsort.so
	$(CXX) -O2 -std=c++17 -Isrc -Iinclude -o $@ $< -L$(LIBDIR) -lmsufsort -lb200sa -Wl,-rpath,'$$ORIGIN'

facade_bench: $(LIBDIR)/facade_bench
$(LIBDIR)/facade_bench: tools/facade_bench.cpp $(LIBDIR)/libmsufsort.so
	$(CXX) -O2 -std=c++17 -Isrc -Iinclude -o $@ $< -L$(LIBDIR) -lmsufsort -lb200sa -Wl,-rpath,'$$ORIGIN'

oracle:
	$(MAKE) -C oracle

emu: tests/emu/libb200sa_emu.so
tests/emu/libb200sa_emu.so: $(KSRC) tests/emu/cuda_emu.h
	$(CXX) -O2 -g -std=c++17 -fPIC -shared -DB200SA_EMU -x c++ -Itests/emu -I$(CSRC) -Wno-unused-function -o $@ $(CSRC)/b200sa.cu

# AddressSanitizer build of the emulator library (tools/emu_asan.sh runs the CPU tier's emulator tests over it): device
# allocations have exact sizes and reused buffers poison the bytes past their current logical size
emu-asan: tests/emu/asan/libb200sa_emu.so
tests/emu/asan/libb200sa_emu.so: $(KSRC) tests/emu/cuda_emu.h
	@mkdir -p tests/emu/asan
	$(CXX) -O1 -g -fno-omit-frame-pointer -fsanitize=address -std=c++17 -fPIC -shared -DB200SA_EMU -DB200SA_EMU_ASAN -x c++ -Itests/emu -I$(CSRC) -Wno-unused-function -o $@ $(CSRC)/b200sa.cu

# UndefinedBehaviorSanitizer build: shifts by the operand width or more (defined differently on the GPU), signed overflow,
# misaligned vector accesses (a fault on the GPU)
emu-ubsan: tests/emu/ubsan/libb200sa_emu.so
tests/emu/ubsan/libb200sa_emu.so: $(KSRC) tests/emu/cuda_emu.h
	@mkdir -p tests/emu/ubsan
	$(CXX) -O1 -g -fno-omit-frame-pointer -fsanitize=undefined -fno-sanitize=vptr -fno-sanitize-recover=undefined -std=c++17 -fPIC -shared -DB200SA_EMU -x c++ -Itests/emu -I$(CSRC) -Wno-unused-function -o $@ $(CSRC)/b200sa.cu

clean:
	rm -rf tests/emu/asan tests/emu/ubsan
	rm -f $(LIBDIR)/*.so $(LIBDIR)/msufsort $(LIBDIR)/facade_bench $(LIBDIR)/ptxas.log tests/emu/*.so
	$(MAKE) -C oracle clean

.PHONY: all lib textgen facade cli facade_bench oracle emu emu-asan emu-ubsan clean
