# Convenience build for C/C++ users (the Python route is `python -m curve25519_b200.build`; both produce the same .so).
NVCC    ?= /usr/local/cuda/bin/nvcc
ARCH    := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2,-Wall --expt-relaxed-constexpr
CSRC    := curve25519_b200/csrc
UNITS   := engine x25519_kernels ed25519_kernels modl_kernels legacy_internals test_kernels comb_table
OBJS    := $(UNITS:%=$(CSRC)/_obj/%.o)
LIB     := curve25519_b200/libcurve25519_b200.so
RPATH   := -Wl,-rpath,'$$ORIGIN/../curve25519_b200'

all: $(LIB)

$(CSRC)/comb_table.cu $(CSRC)/curve_constants.cuh: tools/gen_base_table.py
	python tools/gen_base_table.py

$(CSRC)/_obj/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) $(CSRC)/kernels.h include/c25519_b200.h include/c25519_legacy.h
	@mkdir -p $(CSRC)/_obj
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) -shared -o $@ $(OBJS) $(ARCH) -Xlinker -Bsymbolic -ldl

# plain-C callers of the C ABI (no Python): legacy wrappers + a host-pointer batch; the sharded entry points under any launcher
c_smoke: $(LIB) tools/c_smoke.c
	gcc -O2 -Iinclude tools/c_smoke.c -Lcurve25519_b200 -lcurve25519_b200 $(RPATH) -o tools/c_smoke

c_sharded: $(LIB) tools/c_sharded.c
	gcc -O2 -Iinclude -I/usr/local/cuda/include tools/c_sharded.c -Lcurve25519_b200 -lcurve25519_b200 -L/usr/local/cuda/lib64 -lcudart $(RPATH) -o tools/c_sharded

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf $(CSRC)/_obj $(LIB) tools/c_smoke tools/c_sharded
.PHONY: all c_smoke c_sharded oracle clean
