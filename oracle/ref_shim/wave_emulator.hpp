// wave_emulator.hpp -- TEST INFRASTRUCTURE ONLY (part of the oracle/_ref recipe, see oracle/build_ref.sh).
//
// Lets the reference's Sorting/*.compute kernels -- thread groups of 1024 with group barriers and wave32 intrinsics
// (WavePrefixCountBits, WavePrefixSum) -- run on a CPU exactly as written: every thread of a group is a stackful
// coroutine, and a scheduler advances them the way the hardware's lock-step waves and barriers would:
//   * lanes of a wave run in lane order until each blocks in a wave intrinsic, blocks at a group barrier, or returns;
//   * when no lane of the wave can run, the lanes waiting in a wave intrinsic ARE its participants (HLSL wave ops act
//     on the active lanes): their results are computed in lane order and they resume;
//   * when every thread of the group waits at a group barrier (or has returned), the barrier opens.
// Groups run one after the other (the kernels only communicate across groups through buffers between dispatches).
// The context switch is a dozen x86-64 instructions (callee-saved registers + stack pointer).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <vector>

#if !defined(__x86_64__)
#error "the wave emulator's context switch is written for x86-64 (the build container); oracle/_ref is optional elsewhere"
#endif

extern "C" void usrt_ref_ctx_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl usrt_ref_ctx_switch
.type usrt_ref_ctx_switch,@function
usrt_ref_ctx_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size usrt_ref_ctx_switch,.-usrt_ref_ctx_switch
)");

namespace wave_emu {

constexpr int kWave = 32;                      // Constants.cginc:5 WARP_SIZE; README.md:9 (wave32 only)
enum State { RUNNABLE, WAIT_WAVE, WAIT_GROUP, DONE };
enum WaveOp { PREFIX_COUNT_BITS = 1, PREFIX_SUM = 2 };

struct Fiber {
    void* sp = nullptr;
    State state = DONE;
    int op = 0;
    uint32_t value = 0, result = 0;
};

static std::vector<Fiber> g_fibers;
static int g_current = -1, g_group = 0;
static void* g_scheduler_sp = nullptr;
static void (*g_kernel)(uint32_t thread_in_group, uint32_t group) = nullptr;
static char* g_stacks = nullptr;
static size_t g_stack_count = 0;
constexpr size_t kStackBytes = 32 * 1024;

inline void to_scheduler() { usrt_ref_ctx_switch(&g_fibers[g_current].sp, g_scheduler_sp); }

// what the kernels call
inline void group_barrier() { g_fibers[g_current].state = WAIT_GROUP; to_scheduler(); }
inline uint32_t wave_op(int op, uint32_t v) {
    Fiber& f = g_fibers[g_current];
    f.state = WAIT_WAVE; f.op = op; f.value = v;
    to_scheduler();
    return g_fibers[g_current].result;
}

static void fiber_entry() {
    g_kernel((uint32_t)g_current, (uint32_t)g_group);
    g_fibers[g_current].state = DONE;
    to_scheduler();
    abort();                                    // a finished fiber is never resumed
}

inline void prepare(Fiber& f, int index) {
    char* top = g_stacks + (size_t)(index + 1) * kStackBytes;       // 16-byte aligned
    void** sp = reinterpret_cast<void**>(top);
    *--sp = nullptr;                                                 // where fiber_entry's caller's return address would be
    *--sp = reinterpret_cast<void*>(&fiber_entry);                   // `ret` of the first switch jumps here
    for (int i = 0; i < 6; ++i) *--sp = nullptr;                     // rbp rbx r12 r13 r14 r15
    f.sp = sp; f.state = RUNNABLE; f.op = 0; f.value = f.result = 0;
}

inline void resume(int index) {
    g_current = index;
    usrt_ref_ctx_switch(&g_scheduler_sp, g_fibers[index].sp);
}

// One thread group of `threads` threads of kernel `k`, group id `group`.
inline void run_group(void (*k)(uint32_t, uint32_t), int threads, int group) {
    if ((size_t)threads > g_stack_count) {
        free(g_stacks);
        g_stacks = static_cast<char*>(aligned_alloc(64, (size_t)threads * kStackBytes));
        g_stack_count = threads;
    }
    g_fibers.assign(threads, Fiber());
    g_kernel = k; g_group = group;
    for (int i = 0; i < threads; ++i) prepare(g_fibers[i], i);
    const int waves = (threads + kWave - 1) / kWave;
    for (;;) {
        for (int w = 0; w < waves; ++w) {
            const int lo = w * kWave, hi = lo + kWave < threads ? lo + kWave : threads;
            for (;;) {
                bool ran = false;
                for (int i = lo; i < hi; ++i)
                    if (g_fibers[i].state == RUNNABLE) { resume(i); ran = true; }
                if (ran) continue;
                // nobody can run: the lanes inside a wave intrinsic are its participants
                int op = 0; uint32_t running = 0; bool any = false;
                for (int i = lo; i < hi; ++i) {
                    Fiber& f = g_fibers[i];
                    if (f.state != WAIT_WAVE) continue;
                    if (any && f.op != op) abort();                  // divergent wave intrinsics: not in these kernels
                    any = true; op = f.op;
                    f.result = running;                              // exclusive prefix over the participating lanes
                    running += (op == PREFIX_COUNT_BITS) ? (f.value ? 1u : 0u) : f.value;
                    f.state = RUNNABLE;
                }
                if (!any) break;                                     // the wave is at a group barrier or finished
            }
        }
        bool waiting = false;
        for (Fiber& f : g_fibers)
            if (f.state == WAIT_GROUP) { f.state = RUNNABLE; waiting = true; }
        if (!waiting) break;                                         // every thread returned
    }
}

inline void dispatch(void (*k)(uint32_t, uint32_t), int threads, int groups) {
    for (int g = 0; g < groups; ++g) run_group(k, threads, g);
}

}  // namespace wave_emu

// ---- the HLSL spellings ----------------------------------------------------------------------------------------------
inline void GroupMemoryBarrierWithGroupSync() { wave_emu::group_barrier(); }
inline void AllMemoryBarrierWithGroupSync() { wave_emu::group_barrier(); }
inline void GroupMemoryBarrier() {}                                   // memory ordering only: nothing to do for coroutines
inline uint32_t WavePrefixCountBits(bool b) { return wave_emu::wave_op(wave_emu::PREFIX_COUNT_BITS, b ? 1u : 0u); }
inline uint32_t WavePrefixSum(uint32_t v) { return wave_emu::wave_op(wave_emu::PREFIX_SUM, v); }
#define groupshared static
