# Builds libreef_b200.so (sm_100a only) and the oracle's C restatement.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unknown-pragmas --expt-relaxed-constexpr
CSRC      := reef_b200/csrc
OBJDIR    := build
SRCS      := $(CSRC)/api.cu $(CSRC)/poseidon.cu $(CSRC)/poseidon_ro.cu $(CSRC)/mle.cu $(CSRC)/msm.cu $(CSRC)/sumcheck.cu $(CSRC)/p2p.cu $(CSRC)/cmt.cu
OBJS      := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(SRCS))
HDRS      := $(wildcard $(CSRC)/*.cuh $(CSRC)/*.h $(CSRC)/*.inc include/*.h)
LIB       := reef_b200/libreef_b200.so

all: $(LIB) oracle

$(OBJDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS)

oracle:
	@if [ -f oracle/c/Makefile ]; then $(MAKE) -C oracle/c; fi

gen:
	python tools/gen_fp_asm.py > $(CSRC)/fp_asm.inc
	python tools/gen_fp_asm.py --consts $(CSRC)/fp_consts.inc > /dev/null
	python tools/gen_poseidon_consts.py $(CSRC)/poseidon_consts.inc

clean:
	rm -rf $(OBJDIR) $(LIB)

.PHONY: all oracle gen clean
