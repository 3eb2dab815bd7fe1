# Builds the product libraries in-tree:
#   ssim_b200/lib/libssim_cuda.so   C-ABI shim + sm_100a kernels   (include/ssim_cuda.h)
#   ssim_b200/lib/librmgr-ssim.so   the reference's C/C++ API      (include/rmgr/ssim.h, ssim-openmp.h)
#   ssim_b200/lib/libssim_imgio.so  the front end's JPEG reader    (include/ssim_imgio.h; host code)
# `make oracle` builds the CPU checkers (test infrastructure) under oracle/.
NVCC    ?= /usr/local/cuda/bin/nvcc
HOSTCXX ?= /usr/bin/g++
ARCH    := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin $(HOSTCXX) -Iinclude -Issim_b200/csrc
CSRC    := ssim_b200/csrc
LIB     := ssim_b200/lib
OBJ     := build/obj

BIN     := ssim_b200/bin

all: $(LIB)/libssim_cuda.so $(LIB)/librmgr-ssim.so $(LIB)/libssim_imgio.so $(BIN)/rmgr-ssim $(BIN)/latency_client

$(OBJ)/%.o: $(CSRC)/%.cu $(CSRC)/ssim_kernels.h $(CSRC)/synth.h include/ssim_cuda.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(OBJ)/%.o: $(CSRC)/%.cpp include/ssim_cuda.h include/rmgr/ssim.h
	@mkdir -p $(OBJ)
	$(HOSTCXX) -O2 -std=c++17 -fPIC -DNDEBUG -Iinclude -I/usr/local/cuda/include -c $< -o $@

$(LIB)/libssim_cuda.so: $(OBJ)/ssim_kernels.o $(OBJ)/ssim_cuda.o
	@mkdir -p $(LIB)
	$(NVCC) $(ARCH) -shared -ccbin $(HOSTCXX) -o $@ $^ -cudart static -ldl -lpthread

$(LIB)/librmgr-ssim.so: $(OBJ)/rmgr_api.o $(LIB)/libssim_cuda.so
	$(HOSTCXX) -shared -o $@ $(OBJ)/rmgr_api.o -L$(LIB) -lssim_cuda -Wl,-rpath,'$$ORIGIN'

$(LIB)/libssim_imgio.so: $(CSRC)/imgio.cpp $(CSRC)/jpeg_reader.h include/ssim_imgio.h
	@mkdir -p $(LIB)
	$(HOSTCXX) -O2 -std=c++17 -fPIC -shared -DNDEBUG -Wall -Wextra -Iinclude -o $@ $(CSRC)/imgio.cpp

# rmgr-ssim: the reference's CLI (src/ssim-cli.cpp) re-done on top of the public API; PNG/PNM I/O via zlib
$(BIN)/rmgr-ssim: $(CSRC)/ssim_cli.cpp $(CSRC)/jpeg_reader.h $(LIB)/librmgr-ssim.so $(LIB)/libssim_cuda.so
	@mkdir -p $(BIN)
	$(HOSTCXX) -O2 -std=c++17 -Iinclude -o $@ $(CSRC)/ssim_cli.cpp -L$(LIB) -lrmgr-ssim -lssim_cuda -lz -Wl,-rpath,'$$ORIGIN/../lib'

# latency_client: what one device-pointer call costs a C++ caller (used by bench.py for the single-pair latency)
$(BIN)/latency_client: $(CSRC)/latency_client.cpp $(LIB)/libssim_cuda.so
	@mkdir -p $(BIN)
	$(HOSTCXX) -O2 -std=c++17 -Iinclude -I/usr/local/cuda/include -o $@ $(CSRC)/latency_client.cpp -L$(LIB) -lssim_cuda -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,'$$ORIGIN/../lib' -Wl,-rpath,/usr/local/cuda/lib64

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build $(LIB) $(BIN)

.PHONY: all oracle clean
