#!/bin/bash
# build a variant of liblgarb.so into build/liblgarb_$1.so with extra nvcc flags $2..
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -fmad=false -Xcompiler -fPIC "$@" -shared -o build/liblgarb_$name.so lgar_c_b200/csrc/lgar_kernels.cu lgar_c_b200/csrc/lgar_finish_smem.cu lgar_c_b200/csrc/lgar_engine.cu 2>&1 | grep -E "rror" 
