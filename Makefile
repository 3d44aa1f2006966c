# Builds the product library (sm_100a only) and the CPU oracle (test infrastructure).
#   make            -> kinectpy_b200/libkinectpy_b200.so  +  oracle/_build/libkp_oracle.so
#   make lib / make oracle / make clean
NVCC      ?= nvcc
ORACLE_CC ?= /usr/bin/gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
# -fmad=false: no FMA contraction anywhere, so decision arithmetic matches the oracle (gcc -ffp-contract=off)
NVFLAGS   := $(ARCH) -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v
CSRC      := kinectpy_b200/csrc
SRCS      := $(wildcard $(CSRC)/*.cu)
OBJS      := $(patsubst $(CSRC)/%.cu,build/%.o,$(SRCS))
LIB       := kinectpy_b200/libkinectpy_b200.so
ORACLE    := oracle/_build/libkp_oracle.so

all: lib oracle
lib: $(LIB)
oracle: $(ORACLE)

build/%.o: $(CSRC)/%.cu $(CSRC)/kp_common.cuh $(CSRC)/kp_grid.cuh include/kp_api.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -cudart static

$(ORACLE): oracle/kp_oracle.c
	@mkdir -p oracle/_build
	$(ORACLE_CC) -O2 -fopenmp -ffp-contract=off -fPIC -shared -fvisibility=hidden -Wall -Wextra -o $@ $< -lm

clean:
	rm -rf build $(LIB) oracle/_build

.PHONY: all lib oracle clean
