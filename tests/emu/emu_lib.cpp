// The product's C-ABI translation unit compiled for the CPU kernel-logic emulator (see cuda_emu.h).
// TEST INFRASTRUCTURE ONLY: built into tests/emu/libradfoam_b200_emu.so by tests/emu/build.py.
#include "cuda_emu.h"

namespace emu {
thread_local CtaRun *run = nullptr;
thread_local Fiber *cur = nullptr;

void fiber_entry() {
    (*run->body)();
    thread_exit();
    cur->done = true;
    // returning ends the context: uc_link resumes the scheduler
}
} // namespace emu

#include "../../radfoam_b200/csrc/radfoam_b200.cu"
