# Builds libvimz_gpu.so (sm_100a) in-tree and the CPU oracle.  `make -j` compiles the four curve
# translation units in parallel.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
EXTRA ?=
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Iinclude $(EXTRA)
SRC := vimz_b200/csrc
OBJ ?= build/obj
LIB ?= vimz_b200/libvimz_gpu.so
CUS := capi comm curve_pallas curve_vesta curve_bn254 curve_grumpkin
OBJS := $(addprefix $(OBJ)/,$(addsuffix .o,$(CUS)))
HDRS := $(wildcard $(SRC)/*.cuh) include/vimz_gpu.h

all: $(LIB) oracle

# A/B variants: make OBJ=build/v1 LIB=build/variants/x.so EXTRA="-DVIMZ_COMBINE_MID=64" build/variants/x.so
$(LIB): $(OBJS)
	@mkdir -p $(dir $@)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart -ldl

$(OBJ)/%.o: $(SRC)/%.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $(OBJ)/$*.ptxas.log || (cat $(OBJ)/$*.ptxas.log; exit 1)

oracle:
	$(MAKE) -C oracle

tools/int_peak: tools/int_peak.cu $(SRC)/fp.cuh
	$(NVCC) $(ARCH) -O3 -I$(SRC) -o $@ $< -ldl

clean:
	rm -rf build vimz_b200/libvimz_gpu.so tools/int_peak
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
