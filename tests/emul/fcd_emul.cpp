// TEST INFRASTRUCTURE ONLY: CPU emulation build of the product sources (see fcd_launch.cuh).
//   g++ -O2 -std=c++17 -DFCD_EMULATE -shared -fPIC -I include -I trapped-modes-ltg_b200/csrc \
//       -o tests/emul/libfcd_emul.so tests/emul/fcd_emul.cpp
#ifndef FCD_EMULATE
#error "build with -DFCD_EMULATE"
#endif
#include "fcd_plan.inl"

// emulation-only knob (not part of include/fcd_b200.h): thread order inside a phase
extern "C" void fcd_emul_set_thread_order(int order) { fcd::rt::emu_thread_order() = order; }
