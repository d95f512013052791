# Builds libsz3b200.so (CUDA kernels + host tail + C ABI) for sm_100a, in-tree.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
# -fmad=false: the reference arithmetic has no FMA contraction (SURVEY.md Appendix A / hard part 1)
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-Wall,-Wno-unused-function \
           -Xptxas -v --expt-relaxed-constexpr
SRC := sz3_b200/csrc
OBJDIR := build/obj
LIB := sz3_b200/lib/libsz3b200.so
LIBC := sz3_b200/lib/libSZ3c.so
CU := api.cu pipeline.cu interp_kernels.cu interp_box.cu encode_kernels.cu misc_kernels.cu blockwise.cu lorenzo.cu decompress.cu huffman_decode.cu zhuf_kernels.cu
CPP := huffman_host.cpp stream_host.cpp
OBJS := $(CU:%.cu=$(OBJDIR)/%.o) $(CPP:%.cpp=$(OBJDIR)/%.o)
HDRS := $(wildcard $(SRC)/*.hpp $(SRC)/*.cuh $(SRC)/*.h) include/sz3b.h

all: $(LIB) $(LIBC)

$(OBJDIR)/%.o: $(SRC)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; false)

$(OBJDIR)/%.o: $(SRC)/%.cpp $(HDRS)
	@mkdir -p $(OBJDIR)
	g++ -O2 -std=c++17 -fPIC -ffp-contract=off -Wall -c $< -o $@

$(LIB): $(OBJS)
	@mkdir -p sz3_b200/lib
	$(NVCC) $(ARCH) -shared -Xlinker --no-undefined -o $@ $(OBJS) -l:libzstd.so.1 -lpthread -ldl

# libSZ3c: the reference's C shim (tools/sz3c) rebuilt on the drop-in headers; depends on libsz3b200 at run time
$(LIBC): sz3_b200/sz3c/sz3c.cpp include/sz3c.h $(wildcard include/SZ3/*.hpp include/SZ3/*/*.hpp) $(LIB)
	g++ -O2 -std=c++17 -fPIC -shared -Iinclude sz3_b200/sz3c/sz3c.cpp -o $@ -Lsz3_b200/lib -lsz3b200 -Wl,-rpath,'$$ORIGIN'

clean:
	rm -rf build sz3_b200/lib/*.so

.PHONY: all clean
