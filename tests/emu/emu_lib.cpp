// The product's C-ABI translation unit compiled for the CPU kernel-logic emulator (see cuda_emu.h).
// TEST INFRASTRUCTURE ONLY: built into tests/emu/libradfoam_b200_emu.so by tests/emu/build.py.
#include "cuda_emu.h"

namespace emu {
thread_local CtaRun *run = nullptr;
thread_local Fiber *cur = nullptr;
Counters counters;

void fiber_entry() {
    RFB_EMU_SWITCHED(nullptr, &run->scheduler_stack, &run->scheduler_stack_size);
    (*run->body)();
    thread_exit();
    cur->done = true;
    // leave for good (a null save slot tells the sanitizer that this stack is finished); returning ends the context
    // and uc_link resumes the scheduler
    RFB_EMU_SWITCH_TO(nullptr, run->scheduler_stack, run->scheduler_stack_size);
}
} // namespace emu

#include "../../radfoam_b200/csrc/radfoam_b200.cu"

// Debug accessor for analysis scripts (tests/tools/tape_stats.py): the walk tape of the last recording forward.
extern "C" int rfe_tape_buffers(rfb_pipeline *p, const void **pool, const void **table, const void **per_ray,
                                uint32_t *table_stride, uint32_t *capacity_chunks) {
    if (!p || !p->tape_valid)
        return 1;
    *pool = p->tape_pool.ptr;
    *table = p->tape_table.ptr;
    *per_ray = p->tape_per_ray.ptr;
    *table_stride = p->tape_table_stride;
    *capacity_chunks = p->tape_capacity;
    return 0;
}

// Debug accessor: the replay schedule of the last recording forward (tile step counts and tile order).
extern "C" int rfe_tape_schedule(rfb_pipeline *p, const uint32_t **tile_steps, const uint32_t **order, uint32_t *blocks) {
    if (!p || !p->tape_valid || !p->tape_scheduled)
        return 1;
    *tile_steps = reinterpret_cast<const uint32_t *>(p->tape_sched.ptr);
    *order = *tile_steps + p->tape_blocks;
    *blocks = p->tape_blocks;
    return 0;
}

// Counters of the emulated reductions (16-byte, 8-byte) since the last reset.
extern "C" void rfe_counters(uint64_t *red_v4, uint64_t *red_v2, uint64_t *collectives, int reset) {
    *red_v4 = emu::counters.red_v4.load();
    *red_v2 = emu::counters.red_v2.load();
    *collectives = emu::counters.collectives.load();
    if (reset) {
        emu::counters.red_v4 = 0;
        emu::counters.red_v2 = 0;
        emu::counters.collectives = 0;
    }
}

// NVSwitch multicast stand-in for tests: `world` buffers of `bytes` each; returns the fake base address to pass as the
// multicast pointer (world == 0 clears the group).
extern "C" void *rfe_set_multicast(void *const *copies, uint32_t world, uint64_t bytes) {
    auto &g = emu::multicast();
    static std::vector<char> fake_space;
    g.copies.clear();
    g.fake = nullptr;
    g.bytes = 0;
    if (world == 0)
        return nullptr;
    fake_space.assign(bytes + 64, 0); // only its address range is used
    g.fake = fake_space.data();
    g.bytes = bytes;
    for (uint32_t w = 0; w < world; ++w)
        g.copies.push_back(reinterpret_cast<char *>(copies[w]));
    return g.fake;
}
