# Convenience targets; the library itself is built by wavefunctiondiffusion_b200/build.py (plain nvcc, sm_100a).
PY ?= python
CUDA ?= /usr/local/cuda

.PHONY: lib demo test-cpu test-gpu bench clean
lib:
	$(PY) -m wavefunctiondiffusion_b200.build

demo: lib
	gcc -std=c99 -O2 examples/c_abi_demo.c -Iinclude -I$(CUDA)/include -Lwavefunctiondiffusion_b200 -lwfd_b200 \
	    -L$(CUDA)/lib64 -lcudart -lm -o examples/c_abi_demo
	@echo "run: LD_LIBRARY_PATH=wavefunctiondiffusion_b200:$(CUDA)/lib64 examples/c_abi_demo"

test-cpu: lib
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu: lib
	$(PY) -m pytest tests -x -q -m gpu

bench: lib
	$(PY) bench.py --gpus 1 --steps 20 --warmup 3

clean:
	rm -rf wavefunctiondiffusion_b200/build wavefunctiondiffusion_b200/libwfd_b200.so examples/c_abi_demo
