# Builds the B200-native aligner in-tree (the .so travels to the GPU box with gpurun).
#   make            -> aim_b200/libaim_b200.so, build/host, oracle/libaim_oracle.so
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
CC        ?= gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -O3 -std=c++17 -lineinfo $(ARCH) -Iinclude -Iaim_b200/csrc -Xcompiler -fPIC,-Wall,-Wextra -cudart static $(EXTRA_NVFLAGS)
CSRC      := aim_b200/csrc
OBJDIR    := build/obj
CU_SRCS   := $(CSRC)/aim_wfa.cu $(CSRC)/aim_wfa_sub.cu $(CSRC)/aim_wfa_long.cu $(CSRC)/aim_dp.cu $(CSRC)/aim_dp_fast.cu $(CSRC)/aim_genasm.cu $(CSRC)/aim_dispatch.cu $(CSRC)/aim_file.cu $(CSRC)/aim_filepipe.cu $(CSRC)/aim_peak.cu
CXX_SRCS  := $(CSRC)/aim_host.cpp
OBJS      := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(CU_SRCS)) $(patsubst $(CSRC)/%.cpp,$(OBJDIR)/%.o,$(CXX_SRCS))
HDRS      := include/aim_b200.h $(CSRC)/aim_internal.h $(CSRC)/aim_wfa_common.cuh $(CSRC)/aim_dp_pack2.cuh $(CSRC)/aim_dp_scan.cuh

all: aim_b200/libaim_b200.so aim_b200/libaim_dpu.so build/host build/aim_genpairs build/diag_xfer build/diag_hostfill oracle/libaim_oracle.so

$(OBJDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; false)

$(OBJDIR)/%.o: $(CSRC)/%.cpp $(HDRS)
	@mkdir -p $(OBJDIR)
	$(CXX) -O2 -std=c++17 -fPIC -Wall -Wextra -Iinclude -I$(CSRC) -c $< -o $@

aim_b200/libaim_b200.so: $(OBJS)
	$(NVCC) -shared $(ARCH) -cudart static -o $@ $(OBJS) -lpthread

# UPMEM host-API adapter (include/dpu.h): the reference's host.c files compile unchanged against it (host-only code)
aim_b200/libaim_dpu.so: $(CSRC)/aim_dpu.cpp include/dpu.h include/aim_b200.h aim_b200/libaim_b200.so
	$(CXX) -O2 -std=c++17 -fPIC -shared -Wall -Wextra -Iinclude $(CSRC)/aim_dpu.cpp -o $@ -Laim_b200 -laim_b200 -Wl,-rpath,'$$ORIGIN' -lpthread

build/host: tools/host.cpp aim_b200/libaim_b200.so include/aim_b200.h
	@mkdir -p build
	$(CXX) -O2 -std=c++17 -Wall -Iinclude tools/host.cpp -o $@ -Laim_b200 -laim_b200 -Wl,-rpath,'$$ORIGIN/../aim_b200' -lpthread -ldl

# copy-only host<->device ceiling for 1..N GPUs
build/diag_xfer: tools/diag_xfer.cu
	@mkdir -p build
	$(NVCC) -O2 -std=c++17 $(ARCH) -cudart static tools/diag_xfer.cu -o $@ -lpthread

# how fast host threads rebuild op rows beside a saturating H2D copy
build/diag_hostfill: tools/diag_hostfill.cu
	@mkdir -p build
	$(NVCC) -O2 -std=c++17 $(ARCH) -cudart static tools/diag_hostfill.cu -o $@ -lpthread

# stand-alone pair-file generator for bench.py's reference arm (host-only code, no CUDA, no libaim_b200.so)
build/aim_genpairs: tools/genpairs.cpp $(CSRC)/aim_host.cpp $(HDRS)
	@mkdir -p build
	$(CXX) -O2 -std=c++17 -Wall -Iinclude -I$(CSRC) tools/genpairs.cpp $(CSRC)/aim_host.cpp -o $@ -lpthread

oracle/libaim_oracle.so: oracle/aim_oracle.c
	$(CC) -O2 -std=gnu11 -fPIC -shared -Wall -o $@ $< -lpthread

clean:
	rm -rf build aim_b200/libaim_b200.so aim_b200/libaim_dpu.so oracle/libaim_oracle.so

.PHONY: all clean
