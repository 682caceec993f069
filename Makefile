# Top-level build: libptb200.so (CUDA product), libsmallpt_host.so + smallpt (C++ host surface), oracle (checker).
# nvcc cross-compiles sm_100a without a GPU.  Built artefacts stay in-tree (git-ignored, shipped by gpurun).
NVCC     ?= /usr/local/cuda/bin/nvcc
CXX      := g++
PKG      := small-pathtracer_b200
CSRC     := $(PKG)/csrc
HOST     := $(PKG)/host
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude -I$(CSRC) -Xptxas -v
OBJDIR   := build

all: $(PKG)/libptb200.so $(PKG)/libsmallpt_host.so $(PKG)/smallpt oracle

$(OBJDIR):
	mkdir -p $(OBJDIR)

# FP64 validation engine: NO FMA contraction (bit parity with the -ffp-contract=off oracle)
$(OBJDIR)/pt_validate.o: $(CSRC)/pt_validate.cu $(CSRC)/pt_scene_dev.h $(CSRC)/pt_internal.h include/ptb200.h include/ptb200_detmath.h | $(OBJDIR)
	$(NVCC) $(NVFLAGS) -fmad=false -c $< -o $@ 2> $(OBJDIR)/pt_validate.ptxas.log || (cat $(OBJDIR)/pt_validate.ptxas.log; false)

# FP32 production engine: no implicit FMA contraction either (every FMA is an explicit fmaf), so that this build and the NVRTC one agree bit for bit
$(OBJDIR)/pt_wavefront.o: $(CSRC)/pt_wavefront.cu $(CSRC)/pt_kernel.cuh $(CSRC)/pt_scene_dev.h $(CSRC)/pt_internal.h $(CSRC)/pt_rng.cuh include/ptb200.h | $(OBJDIR)
	$(NVCC) $(NVFLAGS) -fmad=false -c $< -o $@ 2> $(OBJDIR)/pt_wavefront.ptxas.log || (cat $(OBJDIR)/pt_wavefront.ptxas.log; false)

$(OBJDIR)/pt_kernel_src.h: $(CSRC)/pt_scene_dev.h $(CSRC)/pt_rng.cuh $(CSRC)/pt_kernel.cuh tools/embed_kernel_src.py | $(OBJDIR)
	python3 tools/embed_kernel_src.py $@ $(CSRC)/pt_scene_dev.h $(CSRC)/pt_rng.cuh $(CSRC)/pt_kernel.cuh

# NVRTC front end for scene-specialised kernels (libnvrtc is dlopen'ed at run time: only its header is needed here)
$(OBJDIR)/pt_jit.o: $(CSRC)/pt_jit.cu $(OBJDIR)/pt_kernel_src.h $(CSRC)/pt_internal.h $(CSRC)/pt_scene_dev.h include/ptb200.h | $(OBJDIR)
	$(NVCC) $(NVFLAGS) -I$(OBJDIR) -c $< -o $@ 2> $(OBJDIR)/pt_jit.ptxas.log || (cat $(OBJDIR)/pt_jit.ptxas.log; false)

$(OBJDIR)/pt_api.o: $(CSRC)/pt_api.cu $(CSRC)/pt_scene_dev.h $(CSRC)/pt_internal.h include/ptb200.h | $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJDIR)/pt_api.ptxas.log || (cat $(OBJDIR)/pt_api.ptxas.log; false)

$(PKG)/libptb200.so: $(OBJDIR)/pt_validate.o $(OBJDIR)/pt_wavefront.o $(OBJDIR)/pt_api.o $(OBJDIR)/pt_jit.o
	$(NVCC) $(ARCH) -shared -cudart static $^ -ldl -o $@

$(PKG)/libsmallpt_host.so: $(HOST)/scenes.cpp $(HOST)/host_capi.cpp $(HOST)/smallpt_b200.hpp $(HOST)/scene_io.hpp include/ptb200.h
	$(CXX) -O2 -std=c++17 -fPIC -shared -Iinclude $(HOST)/scenes.cpp $(HOST)/host_capi.cpp -o $@

$(PKG)/smallpt: $(HOST)/smallpt_main.cpp $(HOST)/scenes.cpp $(HOST)/smallpt_b200.hpp $(HOST)/scene_io.hpp $(PKG)/libptb200.so
	$(CXX) -O2 -std=c++17 -Iinclude $(HOST)/smallpt_main.cpp $(HOST)/scenes.cpp -L$(PKG) -lptb200 -Wl,-rpath,'$$ORIGIN' -o $@

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf $(OBJDIR) $(PKG)/libptb200.so $(PKG)/libsmallpt_host.so $(PKG)/smallpt
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
