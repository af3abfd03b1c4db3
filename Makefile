# Builds libccx.so (sm_100a only) in-tree; `python -c "import __graft_entry__ as g; g.build()"` calls this.
NVCC ?= /usr/local/cuda/bin/nvcc
PKG := imagecaptioningconvnext_b200
SRC := $(wildcard $(PKG)/csrc/*.cu)
HDR := $(wildcard $(PKG)/csrc/*.h $(PKG)/csrc/*.cuh include/*.h)
OBJ := $(patsubst $(PKG)/csrc/%.cu,build/%.o,$(SRC))
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden

all: $(PKG)/libccx.so

build/%.o: $(PKG)/csrc/%.cu $(HDR)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(PKG)/libccx.so: $(OBJ)
	$(NVCC) -shared -o $@ $(OBJ) -lcudart -Wno-deprecated-gpu-targets

GEMM_SRC := $(PKG)/csrc/gemm_tcgen05.cu $(PKG)/csrc/gemm_tcgen05_2cta.cu $(PKG)/csrc/gemm_skinny.cu $(PKG)/csrc/prof.cu
tools/gemm_selftest: tools/gemm_selftest.cu $(GEMM_SRC) $(HDR)
	$(NVCC) $(NVFLAGS) -I$(PKG)/csrc -Iinclude -o $@ tools/gemm_selftest.cu $(GEMM_SRC)

clean:
	rm -rf build $(PKG)/libccx.so
