# Builds libtmvb.so (sm_100a) without Python -- what a maintainer of TopicModelsVB.jl needs next to julia/*.jl.
# Same flags as topicmodelsvb.jl_b200/_lib.py:build() (which __graft_entry__.build() uses).
NVCC  ?= /usr/local/cuda/bin/nvcc
FLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
SRC   := $(wildcard topicmodelsvb.jl_b200/csrc/*.cu)
OBJ   := $(SRC:.cu=.o)
LIB   := topicmodelsvb.jl_b200/libtmvb.so

all: $(LIB)

%.o: %.cu $(wildcard topicmodelsvb.jl_b200/csrc/*.cuh) include/tmvb.h
	$(NVCC) $(FLAGS) -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) $(FLAGS) -shared $(OBJ) -o $@

oracle:            # test infrastructure only
	$(MAKE) -C oracle

test-cpu: all oracle
	python -m pytest tests -q -m "not gpu"

clean:
	rm -f $(OBJ) $(LIB)

.PHONY: all oracle test-cpu clean
