# Builds fei_b200/libfeiscan.so (sm_100a only) and the oracle's C restatement.
# `python -c "import __graft_entry__ as g; g.build()"` drives this file.
NVCC      ?= nvcc
CXX       ?= g++
CC        ?= gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v --expt-relaxed-constexpr
CXXFLAGS  := -O2 -std=c++17 -fPIC -Wall -I/usr/local/cuda/include
SRC       := fei_b200/csrc
OBJ       := build/obj
CU_SRCS   := $(wildcard $(SRC)/*.cu)
CPP_SRCS  := $(wildcard $(SRC)/*.cpp)
OBJS      := $(patsubst $(SRC)/%.cu,$(OBJ)/%.cu.o,$(CU_SRCS)) $(patsubst $(SRC)/%.cpp,$(OBJ)/%.cpp.o,$(CPP_SRCS))
HDRS      := $(wildcard $(SRC)/*.h) $(wildcard $(SRC)/*.cuh) $(wildcard include/*.h)
LIB       := fei_b200/libfeiscan.so
ORACLE_LIB := oracle/libchain_oracle.so
FASTCOLS  := fei_b200/_fastcols.so
PY_INC    := $(shell python3 -c "import sysconfig; print('-I' + sysconfig.get_paths()['include'])")

all: $(LIB) $(ORACLE_LIB) $(FASTCOLS)

$(OBJ)/%.cu.o: $(SRC)/%.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJ)/$*.ptxas.log || (cat $(OBJ)/$*.ptxas.log; exit 1)

$(OBJ)/%.cpp.o: $(SRC)/%.cpp $(HDRS)
	@mkdir -p $(OBJ)
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -ldl -lpthread

$(ORACLE_LIB): oracle/chain_oracle.c
	$(CC) -O2 -fPIC -shared -Wall -o $@ $<

# CPython helper for the Python host layer (attribute marshalling of block objects); host glue, no compute
$(FASTCOLS): fei_b200/_fastcols.c
	$(CC) -O2 -fPIC -shared -Wall $(PY_INC) -o $@ $< || echo "warning: _fastcols not built (no Python headers?): the pure-Python marshalling path will be used"

clean:
	rm -rf build $(LIB) $(ORACLE_LIB) $(FASTCOLS)

.PHONY: all clean
