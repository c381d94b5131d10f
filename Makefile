# Builds libalm_ocr.so for sm_100a (nvcc cross-compiles without a GPU).
NVCC ?= /usr/local/cuda/bin/nvcc
SRC_DIR := advancedliteratemachinery_b200/csrc
OUT := advancedliteratemachinery_b200/libalm_ocr.so
SRCS := $(wildcard $(SRC_DIR)/*.cu)
OBJS := $(patsubst $(SRC_DIR)/%.cu,build/%.o,$(SRCS))
HDRS := $(wildcard $(SRC_DIR)/*.h) $(wildcard $(SRC_DIR)/*.cuh) include/alm_ocr.h
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden \
           -Xptxas -v --expt-relaxed-constexpr

all: $(OUT)

build/%.o: $(SRC_DIR)/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(OUT): $(OBJS)
	$(NVCC) -shared -gencode arch=compute_100a,code=sm_100a -o $@ $(OBJS) -cudart static -ldl

clean:
	rm -rf build $(OUT)
.PHONY: all clean
