# Builds libfv2d_b200.so (CUDA kernels + C ABI, sm_100a only), the C++ host driver, and the
# test oracles.  `python -c "import __graft_entry__ as g; g.build()"` runs `make all`.
NVCC     ?= /usr/local/cuda/bin/nvcc
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := -std=c++17 -O3 $(ARCH) -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xcudafe --diag_suppress=177
CSRC     := fv2d_b200/csrc
OBJDIR   := build
LIB      := fv2d_b200/libfv2d_b200.so
HOSTHDR  := $(wildcard fv2d_b200/host/*.h) $(wildcard include/*.h) $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h)

.PHONY: all lib driver oracle ref clean
all: lib driver oracle

lib: $(LIB)
driver: fv2d_b200/fv2d_b200_main

$(OBJDIR)/fv2d_ops.o: $(CSRC)/fv2d_ops.cu $(HOSTHDR)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) --fmad=false -c $< -o $@

$(OBJDIR)/fv2d_stream.o: $(CSRC)/fv2d_stream.cu $(HOSTHDR)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) --fmad=false -c $< -o $@

# the fused sweep: dispatcher + small kernels, and one translation unit per Riemann solver.
# --fmad=false: the kernel spells its fused multiply-adds out (fma()); contraction left to the compiler
# can differ between the copies of the unrolled row loop, which would make a row's last bit depend on
# the work decomposition (N-GPU == 1-GPU bitwise is a contract)
# (0 = HLL, 1 = HLLC, 2 = FSLP) holding that solver's 36 kernel instantiations (parallel under -j)
$(OBJDIR)/fv2d_sweep.o: $(CSRC)/fv2d_sweep.cu $(HOSTHDR)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) --fmad=false $(SWEEPFLAGS) -c $< -o $@

$(OBJDIR)/fv2d_sweep_s%.o: $(CSRC)/fv2d_sweep.cu $(HOSTHDR)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) --fmad=false $(SWEEPFLAGS) -DFV2D_SOLVER_ONLY=$* -Xptxas -v -c $< -o $@ 2> $(OBJDIR)/fv2d_sweep_s$*.ptxas.log || (cat $(OBJDIR)/fv2d_sweep_s$*.ptxas.log; false)

$(OBJDIR)/fv2d_capi.o: $(CSRC)/fv2d_capi.cu $(HOSTHDR)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -Xcompiler -fopenmp -c $< -o $@

SWEEPOBJ := $(OBJDIR)/fv2d_sweep.o $(OBJDIR)/fv2d_sweep_s0.o $(OBJDIR)/fv2d_sweep_s1.o $(OBJDIR)/fv2d_sweep_s2.o
$(LIB): $(OBJDIR)/fv2d_ops.o $(OBJDIR)/fv2d_stream.o $(SWEEPOBJ) $(OBJDIR)/fv2d_capi.o
	$(NVCC) $(ARCH) -shared -o $@ $^ -cudart static -Xcompiler -fopenmp

fv2d_b200/fv2d_b200_main: fv2d_b200/host/main.cpp $(HOSTHDR) $(LIB)
	/usr/bin/g++ -std=c++17 -O2 -Wall -fopenmp -Iinclude fv2d_b200/host/main.cpp -o $@ -Lfv2d_b200 -lfv2d_b200 -Wl,-rpath,'$$ORIGIN'

oracle:
	$(MAKE) -C oracle oracle

ref:
	$(MAKE) -C oracle ref

clean:
	rm -rf $(OBJDIR) $(LIB) fv2d_b200/fv2d_b200_main
	$(MAKE) -C oracle clean
