# Builds libmorsi_cuda (sm_100a kernels + C ABI), libmorsi_compat (reference
# signatures) and the `morsi` host program.  nvcc cross-compiles without a GPU.
#   make            lib + compat (+ morsi CLI when the reference iio.c is reachable)
#   make oracle     the CPU checker under oracle/ (test infrastructure)
NVCC   ?= nvcc
CC     ?= gcc
REF    ?= /root/reference
ARCH   := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v $(EXTRA)
SRC    := imscript_b200/csrc
OUT    := imscript_b200/lib
OBJ    := build/obj

CU_SRCS := $(wildcard $(SRC)/*.cu)
CU_OBJS := $(patsubst $(SRC)/%.cu,$(OBJ)/%.o,$(CU_SRCS)) $(OBJ)/k_disk_p1.o $(OBJ)/k_disk_p2.o $(OBJ)/k_disk_p3.o
HDRS    := $(wildcard $(SRC)/*.cuh) $(wildcard $(SRC)/*.h) include/morsi_cuda.h

all: $(OUT)/libmorsi_cuda.so $(OUT)/libmorsi_compat.so cli $(OUT)/compat_caller

$(OBJ)/%.o: $(SRC)/%.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJ)/$*.ptxas.log || (cat $(OBJ)/$*.ptxas.log; false)

# k_disk.cu holds 17 shapes x 12 kernels: compiled as four units so that -j parallelises
$(OBJ)/k_disk_p%.o: $(SRC)/k_disk.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -DMORSI_DISK_PART=$* -c $< -o $@ 2> $(OBJ)/k_disk_p$*.ptxas.log || (cat $(OBJ)/k_disk_p$*.ptxas.log; false)

$(OBJ)/element.o: $(SRC)/element.c $(HDRS)
	@mkdir -p $(OBJ)
	$(CC) -O2 -fPIC -Wall -c $< -o $@

$(OUT)/libmorsi_cuda.so: $(CU_OBJS) $(OBJ)/element.o
	@mkdir -p $(OUT)
	$(NVCC) $(ARCH) -shared -o $@ $^ -lm

$(OUT)/libmorsi_compat.so: $(SRC)/compat.c $(OUT)/libmorsi_cuda.so
	$(CC) -O2 -fPIC -Wall -shared -o $@ $(SRC)/compat.c -L$(OUT) -lmorsi_cuda -Wl,-rpath,'$$ORIGIN'

# a library-style caller of the reference API (tests/c/compat_caller.c), linked the way
# INTEGRATION.md tells a maintainer to relink corrview.c: -lmorsi_compat -lmorsi_cuda
$(OUT)/compat_caller: tests/c/compat_caller.c $(OUT)/libmorsi_compat.so
	$(CC) -O2 -Wall -o $@ tests/c/compat_caller.c -L$(OUT) -lmorsi_compat -lmorsi_cuda -Wl,-rpath,'$$ORIGIN'

# The CLI keeps the reference's image I/O: iio.c is compiled from the reference
# tree (never copied).  Where the tree is absent the prebuilt binary is kept.
ifneq ($(wildcard $(REF)/src/iio.c),)
cli: $(OUT)/morsi $(OUT)/im_like
# the multi-call form of src/im.c:4-6: morsi_main.c with its main hidden
$(OUT)/im_like: tests/c/im_like.c $(SRC)/morsi_main.c $(OBJ)/iio.o $(OUT)/libmorsi_cuda.so
	$(CC) -O2 -Wall -DHIDE_ALL_MAINS -o $@ tests/c/im_like.c $(SRC)/morsi_main.c $(OBJ)/iio.o -L$(OUT) -lmorsi_cuda -lm -Wl,-rpath,'$$ORIGIN'
$(OBJ)/iio.o: $(REF)/src/iio.c
	@mkdir -p $(OBJ)
	$(CC) -O3 -w -c $< -o $@
$(OUT)/morsi: $(SRC)/morsi_main.c $(OBJ)/iio.o $(OUT)/libmorsi_cuda.so
	$(CC) -O2 -Wall -o $@ $(SRC)/morsi_main.c $(OBJ)/iio.o -L$(OUT) -lmorsi_cuda -lm -Wl,-rpath,'$$ORIGIN'
else
cli:
	@echo "morsi CLI: $(REF)/src/iio.c absent, keeping prebuilt $(OUT)/morsi"
endif

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build $(OUT)

.PHONY: all cli oracle clean
