# Plain build of the C-ABI library and its C++ drivers for integrators who do not go through Python
# (`python -c "import __graft_entry__ as g; g.build()"` does the same, incrementally, plus the test oracle).
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
NVCCFLAGS ?= -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC
CSRC      := halo2_snark_aggregator_b200/csrc
SOURCES   := capi ntt msm synth witness witness_recorder poly scan sort arguments quotient codec
OBJS      := $(addprefix $(CSRC)/build/,$(addsuffix .o,$(SOURCES)))
LIB       := halo2_snark_aggregator_b200/libh2agg.so
HEADERS   := $(wildcard $(CSRC)/*.cuh $(CSRC)/*.hpp $(CSRC)/*.h) $(CSRC)/gen/mont_mul_bn254.inc include/h2agg.h

all: $(LIB) tests/cpp/prover_main tests/cpp/witness_main

$(CSRC)/build/%.o: $(CSRC)/%.cu $(HEADERS)
	@mkdir -p $(CSRC)/build
	$(NVCC) $(NVCCFLAGS) -c -o $@ $<

$(LIB): $(OBJS)
	$(NVCC) -shared -gencode arch=compute_100a,code=sm_100a -o $@ $(OBJS)

tests/cpp/%: tests/cpp/%.cpp $(LIB) include/h2agg.h include/h2agg.hpp include/h2agg_prover.hpp
	$(CXX) -std=c++17 -O2 -Wall -Wextra -Iinclude $< -o $@ -Lhalo2_snark_aggregator_b200 -lh2agg '-Wl,-rpath,$$ORIGIN/../../halo2_snark_aggregator_b200'

clean:
	rm -rf $(CSRC)/build $(LIB) tests/cpp/prover_main tests/cpp/witness_main

.PHONY: all clean
