// fcd_b200.cu -- the product translation unit: libfcd_b200.so for sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC \
//        -I include -o libfcd_b200.so fcd_b200.cu
#include "fcd_plan.inl"
