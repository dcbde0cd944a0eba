# voldor_b200 build: sm_100a CUDA library (the product) + oracle (test infrastructure)
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 $(ARCH) -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden,-O3,-ffp-contract=off
CSRC := voldor_b200/csrc
SRCS := $(wildcard $(CSRC)/*.cu)
OBJS := $(patsubst $(CSRC)/%.cu,build/%.o,$(SRCS))
HDRS := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) $(wildcard include/*.h)
LIB := voldor_b200/libvoldor_b200.so

.PHONY: all lib oracle clean
all: lib oracle

lib: $(LIB)

build/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -Xlinker -Bsymbolic -o $@ $^ -lcudart

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf build $(LIB)
	$(MAKE) -C oracle clean
