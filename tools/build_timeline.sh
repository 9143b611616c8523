#!/bin/bash
# instrumented debug build (phase stamps in conv_tc_kernel); use with D3FK_LIB=tools/libd3fk_tl.so
cd "$(dirname "$0")/../denoising_diffusion_deep_fake_b200/csrc" && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -diag-suppress 550 -DD3FK_TIMELINE -o ../../tools/libd3fk_tl.so *.cu
