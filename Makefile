# Build of the B200-native P3DFFT++ transform path.
#   make            -> p3dfft.3_b200/lib/libp3dfft.3.so   (host C++ + sm_100a CUDA layer; the product)
#   make emu        -> tools/cuda_emu/_build/libp3dfft_emu.so (CPU-thread emulation of the kernels; dev/test tool only)
#   make -C samples -> the reference's own sample programs, compiled UNMODIFIED from REFERENCE against this library
#   make oracle     -> oracle/_build (C restatement) and, when REFERENCE exists, oracle/_ref (reference host code on shims)
PKG      := p3dfft.3_b200
NVCC     := nvcc
CXX      := g++
ARCH     := -gencode arch=compute_100a,code=sm_100a
# MPI=real : build against a real MPI (mpi.h / libmpi from MPI_INC / MPI_LIB, e.g. `mpicxx -show`) instead of the one-host
#            mini-MPI of this repository (include/compat/mpi.h + host/minimpi.cpp, which would otherwise export MPI_* symbols
#            with MPI_Comm = int and clash with the application's MPI).  Default: the mini-MPI (this image has no MPI).
MPI      ?= mini
ifeq ($(MPI),real)
MPIINC   := $(MPI_INC)
MPILIB   := $(MPI_LIB)
MINIMPI  :=
else
MPIINC   := -Iinclude/compat
MPILIB   :=
MINIMPI  := minimpi
endif
CXXFLAGS := -O2 -std=c++17 -fPIC -fno-gnu-unique -Wall -Wno-comment -Iinclude $(MPIINC) -I$(PKG)/host
NVFLAGS  := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fno-gnu-unique -Iinclude -I$(PKG)/csrc $(EXTRA_NVFLAGS)
HOSTSRC  := geometry registry planner executor cwrap $(MINIMPI)
LIBDIR   ?= $(PKG)/lib
OBJDIR   := $(LIBDIR)/obj
HOSTOBJ  := $(HOSTSRC:%=$(OBJDIR)/%.o)
LIB      := $(LIBDIR)/libp3dfft.3.so
EMUDIR   := tools/cuda_emu/_build
EMULIB   := $(EMUDIR)/libp3dfft_emu.so
CUDA_HOME ?= /usr/local/cuda

all: $(LIB)

$(OBJDIR)/%.o: $(PKG)/host/%.cpp include/p3dfft.h include/p3dfft_b200.h include/Cwrap.h $(PKG)/host/plan.h
	@mkdir -p $(OBJDIR)
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(OBJDIR)/gpu_layer.o: $(PKG)/csrc/gpu_layer.cu $(wildcard $(PKG)/csrc/*.cuh) include/p3dfft_b200.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@

# pipelined pow2 kernels: one object per (precision, kind); kind 5 = DCT-I on complex data, 13 = every r2r kind (DCT/DST I-IV) with the kind read at run time
PIPEKEYS := 4_1 4_2 4_3 4_4 4_5 4_13 8_1 8_2 8_3 8_4 8_5 8_13
PIPEOBJ  := $(PIPEKEYS:%=$(OBJDIR)/pipe_%.o)
$(OBJDIR)/pipe_%.o: $(PKG)/csrc/pow2_pipe_inst.cu $(wildcard $(PKG)/csrc/*.cuh) include/p3dfft_b200.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -DPIPE_PREC=$(word 1,$(subst _, ,$*)) -DPIPE_KIND=$(word 2,$(subst _, ,$*)) -c $< -o $@

# smooth lengths 3|5|7 x 2^k: one object per (precision, kind)
MIXKEYS := 4_1 4_2 4_3 4_4 8_1 8_2 8_3 8_4
MIXOBJ  := $(MIXKEYS:%=$(OBJDIR)/mixed_%.o)
$(OBJDIR)/mixed_%.o: $(PKG)/csrc/mixed_pipe_inst.cu $(wildcard $(PKG)/csrc/*.cuh) include/p3dfft_b200.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -DPIPE_PREC=$(word 1,$(subst _, ,$*)) -DPIPE_KIND=$(word 2,$(subst _, ,$*)) -c $< -o $@

# tensor-load kernels (strided inputs): one object per (precision, kind 1..3)
TLKEYS := 4_1 4_2 4_3 8_1 8_2 8_3
TLOBJ  := $(TLKEYS:%=$(OBJDIR)/tload_%.o)
$(OBJDIR)/tload_%.o: $(PKG)/csrc/pow2_tload_inst.cu $(wildcard $(PKG)/csrc/*.cuh) include/p3dfft_b200.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -DPIPE_PREC=$(word 1,$(subst _, ,$*)) -DPIPE_KIND=$(word 2,$(subst _, ,$*)) -c $< -o $@

$(OBJDIR)/fastcore_inst.o: $(PKG)/csrc/fastcore_inst.cu $(wildcard $(PKG)/csrc/*.cuh) include/p3dfft_b200.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(HOSTOBJ) $(OBJDIR)/gpu_layer.o $(OBJDIR)/fastcore_inst.o $(PIPEOBJ) $(MIXOBJ) $(TLOBJ)
	$(CXX) -shared -o $@ $^ -L$(CUDA_HOME)/lib64 -lcudart_static $(MPILIB) -ldl -lrt -lpthread

emu: $(EMULIB)
$(EMUDIR)/gpu_layer_emu.o: $(PKG)/csrc/gpu_layer.cu $(wildcard $(PKG)/csrc/*.cuh) tools/cuda_emu/cuda_runtime.h
	@mkdir -p $(EMUDIR)
	$(CXX) -O1 -std=c++17 -fPIC -fno-gnu-unique -x c++ -Itools/cuda_emu -Iinclude -I$(PKG)/csrc -c $< -o $@
$(EMUDIR)/emu_globals.o: tools/cuda_emu/emu_globals.cpp tools/cuda_emu/cuda_runtime.h
	@mkdir -p $(EMUDIR)
	$(CXX) -O1 -std=c++17 -fPIC -fno-gnu-unique -Itools/cuda_emu -c $< -o $@
EMUPIPEOBJ := $(PIPEKEYS:%=$(EMUDIR)/pipe_%.o)
$(EMUDIR)/pipe_%.o: $(PKG)/csrc/pow2_pipe_inst.cu $(wildcard $(PKG)/csrc/*.cuh) tools/cuda_emu/cuda_runtime.h
	@mkdir -p $(EMUDIR)
	$(CXX) -O1 -std=c++17 -fPIC -fno-gnu-unique -x c++ -Itools/cuda_emu -Iinclude -I$(PKG)/csrc -DPIPE_PREC=$(word 1,$(subst _, ,$*)) -DPIPE_KIND=$(word 2,$(subst _, ,$*)) -c $< -o $@
EMUMIXOBJ := $(MIXKEYS:%=$(EMUDIR)/mixed_%.o)
$(EMUDIR)/mixed_%.o: $(PKG)/csrc/mixed_pipe_inst.cu $(wildcard $(PKG)/csrc/*.cuh) tools/cuda_emu/cuda_runtime.h
	@mkdir -p $(EMUDIR)
	$(CXX) -O1 -std=c++17 -fPIC -fno-gnu-unique -x c++ -Itools/cuda_emu -Iinclude -I$(PKG)/csrc -DPIPE_PREC=$(word 1,$(subst _, ,$*)) -DPIPE_KIND=$(word 2,$(subst _, ,$*)) -c $< -o $@
EMUTLOBJ := $(TLKEYS:%=$(EMUDIR)/tload_%.o)
$(EMUDIR)/tload_%.o: $(PKG)/csrc/pow2_tload_inst.cu $(wildcard $(PKG)/csrc/*.cuh) tools/cuda_emu/cuda_runtime.h
	@mkdir -p $(EMUDIR)
	$(CXX) -O1 -std=c++17 -fPIC -fno-gnu-unique -x c++ -Itools/cuda_emu -Iinclude -I$(PKG)/csrc -DPIPE_PREC=$(word 1,$(subst _, ,$*)) -DPIPE_KIND=$(word 2,$(subst _, ,$*)) -c $< -o $@
$(EMUDIR)/fastcore_inst.o: $(PKG)/csrc/fastcore_inst.cu $(wildcard $(PKG)/csrc/*.cuh) tools/cuda_emu/cuda_runtime.h
	@mkdir -p $(EMUDIR)
	$(CXX) -O1 -std=c++17 -fPIC -fno-gnu-unique -x c++ -Itools/cuda_emu -Iinclude -I$(PKG)/csrc -c $< -o $@
$(EMULIB): $(HOSTOBJ) $(EMUDIR)/gpu_layer_emu.o $(EMUDIR)/emu_globals.o $(EMUDIR)/fastcore_inst.o $(EMUPIPEOBJ) $(EMUMIXOBJ) $(EMUTLOBJ)
	$(CXX) -shared -o $@ $^ -lrt -lpthread

clean:
	rm -rf $(PKG)/lib $(EMUDIR)

.PHONY: all emu clean
