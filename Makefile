# Top-level build: CUDA library (sm_100a only), synthetic workload generator, oracle.
NVCC    ?= /usr/local/cuda/bin/nvcc
CC      ?= gcc
ARCH    := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v
CSRC    := msamtools_b200/csrc
LIB     := msamtools_b200/libmsamtools_b200.so
SYNTH   := msamtools_b200/libmsamsynth.so
HOSTLIB := msamtools_b200/libmsamhost.so

CLI     := msamtools_b200/bin/msamtools
HOSTSRC := $(CSRC)/host/bamio.c $(CSRC)/host/finflate.c $(CSRC)/host/crc32x.c $(CSRC)/host/gzpar.c $(CSRC)/host/margs.c $(CSRC)/host/keyorder.c $(CSRC)/host/recwalk.c

all: $(LIB) $(SYNTH) $(HOSTLIB) $(CLI) oracle

$(LIB): $(CSRC)/api.cu $(wildcard $(CSRC)/*.cuh) $(CSRC)/host/recindex.c include/msamtools_b200.h
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)/api.cu -ldl

# CPU-only helpers of the ABI (record index, QNAME-boundary split): what the reference arm of bench.py loads
$(HOSTLIB): $(CSRC)/host/recindex.c
	$(CC) -O2 -g -std=gnu99 -Wall -Wextra -fPIC -shared -o $@ $<

$(SYNTH): $(CSRC)/synth.c
	$(CC) -O2 -g -std=gnu99 -Wall -Wextra -fPIC -shared -o $@ $< -lm

# drop-in CLI: plain C host (own BAM/SAM/BGZF I/O over zlib) linked against the CUDA library
$(CLI): $(CSRC)/cli/msamtools_main.c $(HOSTSRC) $(wildcard $(CSRC)/host/*.h) include/msamtools_b200.h $(LIB)
	mkdir -p msamtools_b200/bin
	$(CC) -O2 -g -std=gnu99 -Wall -Wextra -o $@ $(CSRC)/cli/msamtools_main.c $(HOSTSRC) -Lmsamtools_b200 -lmsamtools_b200 -Wl,-rpath,'$$ORIGIN/..' -lz -lm -lpthread

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(LIB) $(SYNTH) $(HOSTLIB) $(CLI); $(MAKE) -C oracle clean

.PHONY: all oracle clean
